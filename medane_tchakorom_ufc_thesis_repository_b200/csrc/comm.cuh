// comm.cuh — communication back ends of the engine (included by engine.cu; one translation unit).
#pragma once
// ------------------------------------------------------------------------------------------------
// communication back ends
// ------------------------------------------------------------------------------------------------
struct Comm {
  int rank = 0, nranks = 1;
  virtual ~Comm() {}
  // in-place sum over ranks of n doubles in device memory, result on every rank, stream ordered
  virtual int allreduce_sum(double *dbuf, int n, cudaStream_t st) = 0;
  virtual int barrier(cudaStream_t st) = 0;
  // true: every block has its own stream on its own GPU/process, so a stream may wait on a word its neighbour writes
  // (the in-process group of the tests serialises through host threads and keeps the barrier)
  virtual bool neighbour_flags() const { return false; }
};

struct SelfComm : Comm {
  int allreduce_sum(double *, int, cudaStream_t) override { return 0; }
  int barrier(cudaStream_t) override { return 0; }
};

// all blocks in one process, one host thread per block: host-side deterministic reduction
struct LocalShared {
  int n;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  long long gen = 0;
  std::vector<std::vector<double>> slot;
  std::vector<double> result;
  bool aborted = false; // a block thread left its solve with an error: nobody may wait for it any more
  explicit LocalShared(int n_) : n(n_), slot(n_), result(0) {}
  // false = the group was aborted (another block failed); the caller returns an error instead of waiting forever
  bool wait_all() {
    std::unique_lock<std::mutex> lk(mu);
    if (aborted) return false;
    long long g = gen;
    if (++arrived == n) { arrived = 0; gen++; cv.notify_all(); }
    else cv.wait(lk, [&] { return gen != g || aborted; });
    return !aborted;
  }
  void abort() {
    std::lock_guard<std::mutex> lk(mu);
    aborted = true;
    cv.notify_all();
  }
  void reset() { // before a group solve starts its block threads
    std::lock_guard<std::mutex> lk(mu);
    aborted = false; arrived = 0;
  }
};
struct LocalComm : Comm {
  LocalShared *sh;
  std::vector<double> host;
  LocalComm(LocalShared *s, int r) : sh(s) { rank = r; nranks = s->n; }
  int allreduce_sum(double *dbuf, int n, cudaStream_t st) override {
    host.resize(n);
    CK(cudaMemcpyAsync(host.data(), dbuf, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    sh->slot[rank] = host;
    if (!sh->wait_all()) MSP_FAIL("group aborted: another block failed");
    for (int i = 0; i < n; i++) {
      double s = 0.0;
      for (int r = 0; r < nranks; r++) s += sh->slot[r][i]; // rank order: identical on every rank
      host[i] = s;
    }
    if (!sh->wait_all()) MSP_FAIL("group aborted: another block failed");
    CK(cudaMemcpyAsync(dbuf, host.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    return 0;
  }
  int barrier(cudaStream_t st) override {
    CK(cudaStreamSynchronize(st));
    if (!sh->wait_all()) MSP_FAIL("group aborted: another block failed");
    return 0;
  }
};

// one process per GPU: NCCL (resolved at run time so that a process that already loaded torch's
// bundled libnccl.so.2 shares it)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
struct NcclApi {
  void *h = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, void *) = nullptr; // NCCL >= 2.18: the communicator of one Jacobi block
  const char *(*GetErrorString)(int) = nullptr;
  bool load() {
    if (h) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) { h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) return false;
    GetUniqueId = (int (*)(ncclUniqueId *))dlsym(h, "ncclGetUniqueId");
    CommInitRank = (int (*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
    CommDestroy = (int (*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
    AllReduce = (int (*)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
    GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    CommSplit = (int (*)(ncclComm_t, int, int, ncclComm_t *, void *))dlsym(h, "ncclCommSplit");
    return GetUniqueId && CommInitRank && CommDestroy && AllReduce;
  }
};
static NcclApi g_nccl;
struct NcclComm : Comm {
  ncclComm_t comm = nullptr;
  double *scratch = nullptr;
  ~NcclComm() override { if (comm) g_nccl.CommDestroy(comm); if (scratch) cudaFree(scratch); }
  int allreduce_sum(double *dbuf, int n, cudaStream_t st) override {
    int rc = g_nccl.AllReduce(dbuf, dbuf, (size_t)n, /*ncclFloat64*/ 8, /*ncclSum*/ 0, comm, st);
    if (rc) MSP_FAIL(std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
    return 0;
  }
  int barrier(cudaStream_t st) override {
    if (!scratch) { CK(cudaMalloc(&scratch, 64)); CK(cudaMemsetAsync(scratch, 0, 64, st)); }
    return allreduce_sum(scratch, 1, st);
  }
  bool neighbour_flags() const override { return true; }
};

