// comm_local.inl — `zkl_comm` for the proving threads of ONE process (include/zkl_b200.h: zkl_comm_local_*).
// Every rank is a host thread with its own zkl_ctx; the contexts may sit on one device (several shards of a segment on one GPU: the
// single-GPU test of the multi-rank flow) or on several devices of a box (cudaMemcpyAsync with cudaMemcpyDefault between devices is a
// peer copy over NVLink when peer access is enabled, a staged copy otherwise). A collective is: publish my pointer, barrier, copy what I
// need out of the other ranks' buffers on my own stream, barrier. One process per GPU over NCCL (zk-lisp_b200/sharded.py: TorchComm) is
// the multi-process counterpart with the same callbacks.
#include <condition_variable>

struct zkl_local_group {
  int world = 0;
  std::mutex mu; std::condition_variable cv;
  int arrived = 0; uint64_t generation = 0; bool broken = false;
  std::vector<const void *> slot;
  struct member { zkl_local_group *g; int rank; cudaStream_t stream = nullptr; void *scratch[2] = {nullptr, nullptr}; size_t cap[2] = {0, 0}; };
  std::vector<member> members;
  std::vector<zkl_comm> comms;
  // a rank that fails before reaching a collective must not hang the others: waits give up after this long and break the group
  int timeout_s = 600;

  int barrier() {
    std::unique_lock<std::mutex> lk(mu);
    if (broken) return 1;
    const uint64_t gen = generation;
    if (++arrived == world) { arrived = 0; generation++; cv.notify_all(); return 0; }
    if (!cv.wait_for(lk, std::chrono::seconds(timeout_s), [&] { return generation != gen || broken; })) { broken = true; cv.notify_all(); return 1; }
    return broken ? 1 : 0;
  }
  void fail() { std::lock_guard<std::mutex> lk(mu); broken = true; cv.notify_all(); }
};

__global__ void k_add_u64(unsigned long long *__restrict__ acc, const unsigned long long *__restrict__ src, size_t count) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < count) acc[t] += src[t];
}

static int local_stream(zkl_local_group::member *m) {
  if (!m->stream && cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) return 1;
  return 0;
}
static int local_scratch(zkl_local_group::member *m, int which, size_t bytes) {
  if (m->cap[which] >= bytes) return 0;
  if (m->scratch[which]) cudaFree(m->scratch[which]);
  m->scratch[which] = nullptr; m->cap[which] = 0;
  if (cudaMalloc(&m->scratch[which], bytes) != cudaSuccess) return 1;
  m->cap[which] = bytes;
  return 0;
}
// out-of-place move: block s of `recv` <- the bytes rank s published, starting `offset_of(s)` bytes in
template <typename F>
static int local_move(void *user, const void *send, void *recv, size_t bytes, F offset_of) {
  auto *m = (zkl_local_group::member *)user; zkl_local_group *g = m->g;
  if (local_stream(m)) { g->fail(); return 1; }
  g->slot[m->rank] = send;
  if (g->barrier()) return 1;
  bool ok = true;
  for (int s = 0; s < g->world && ok; s++)
    ok = cudaMemcpyAsync((char *)recv + (size_t)s * bytes, (const char *)g->slot[s] + offset_of(s), bytes, cudaMemcpyDefault, m->stream) == cudaSuccess;
  ok = ok && cudaStreamSynchronize(m->stream) == cudaSuccess;
  if (!ok) { g->fail(); return 1; }
  return g->barrier();
}
static int local_all_gather(void *user, const void *send, void *recv, size_t bytes) {
  return local_move(user, send, recv, bytes, [](int) { return (size_t)0; });
}
static int local_all_to_all(void *user, const void *send, void *recv, size_t bytes) {
  const int rank = ((zkl_local_group::member *)user)->rank;
  return local_move(user, send, recv, bytes, [=](int) { return (size_t)rank * bytes; });
}
static int local_all_reduce_sum_u64(void *user, void *buf, size_t count) {
  auto *m = (zkl_local_group::member *)user; zkl_local_group *g = m->g;
  if (g->world == 1) return 0;
  if (local_stream(m) || local_scratch(m, 0, count * 8) || local_scratch(m, 1, count * 8)) { g->fail(); return 1; }
  g->slot[m->rank] = buf;
  if (g->barrier()) return 1;
  auto *acc = (unsigned long long *)m->scratch[0], *tmp = (unsigned long long *)m->scratch[1];
  bool ok = cudaMemcpyAsync(acc, buf, count * 8, cudaMemcpyDefault, m->stream) == cudaSuccess;
  for (int s = 0; s < g->world && ok; s++) {
    if (s == m->rank) continue;
    ok = cudaMemcpyAsync(tmp, g->slot[s], count * 8, cudaMemcpyDefault, m->stream) == cudaSuccess;
    if (ok) { k_add_u64<<<(unsigned)((count + 255) / 256), 256, 0, m->stream>>>(acc, tmp, count); ok = cudaGetLastError() == cudaSuccess; }
  }
  ok = ok && cudaStreamSynchronize(m->stream) == cudaSuccess;
  if (!ok) { g->fail(); return 1; }
  if (g->barrier()) return 1;            // every rank has read every buffer: they may be overwritten now
  ok = cudaMemcpyAsync(buf, acc, count * 8, cudaMemcpyDefault, m->stream) == cudaSuccess && cudaStreamSynchronize(m->stream) == cudaSuccess;
  if (!ok) { g->fail(); return 1; }
  return 0;
}

extern "C" int zkl_comm_local_create(int world, zkl_local_group **out) {
  if (world < 1 || world > 64 || !out) return 1;
  auto *g = new zkl_local_group();
  g->world = world; g->slot.assign(world, nullptr); g->members.resize(world); g->comms.resize(world);
  for (int r = 0; r < world; r++) {
    g->members[r].g = g; g->members[r].rank = r;
    g->comms[r] = zkl_comm{r, world, &g->members[r], local_all_gather, local_all_to_all, local_all_reduce_sum_u64};
  }
  *out = g;
  return 0;
}
extern "C" const zkl_comm *zkl_comm_local_rank(zkl_local_group *g, int rank) {
  return (g && rank >= 0 && rank < g->world) ? &g->comms[rank] : nullptr;
}
extern "C" void zkl_comm_local_destroy(zkl_local_group *g) {
  if (!g) return;
  // streams and scratch belong to the device that was current when the rank first used them; cudaFree / cudaStreamDestroy accept
  // them from any current device of the same process
  for (auto &m : g->members) { for (void *p : m.scratch) if (p) cudaFree(p); if (m.stream) cudaStreamDestroy(m.stream); }
  delete g;
}
