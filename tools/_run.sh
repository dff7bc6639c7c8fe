python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -n 2
