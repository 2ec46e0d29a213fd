// dist.cu -- row-partitioned, iterated SpMV over the GPUs of one NVSwitch box:
//            x_{k+1} = A x_k   (include/spmv_b200.h, "multi-GPU" section).
//
// NEW SURFACE -- the reference is single-GPU.  What it offers is the partition
// rule: contiguous row ranges with balanced nnz, a range closed as soon as its running nnz
// reaches total/parts (partition_csr_rows, reference src/csr.c:218-276).
// spmv_b200_partition_rows restates that rule and rounds cuts up to hack boundaries (32 rows).
//
// Rank r owns rows [r0, r1) of A and the matching slice of x.  Its rows reference the global
// column range [c0, c1) -- own slice plus a halo for a banded / stencil matrix, everything for a
// general one.  The local x buffer covers [c0, c1); the shard's column indices are stored
// relative to c0; x is double-buffered (step k reads X[k%2], writes the own slice of X[(k+1)%2]).
//
// One step, mode PUSH (halo plans, at most two peers per boundary segment):
//     wait_kernel      every neighbour has published epoch >= mine: its pushes into X[src] have
//                      landed and it has finished every row that reads the halo of X[dst]
//     boundary rows    SpMV whose epilogue also stores the rows a neighbour needs straight into
//                      that neighbour's halo (peer HBM over NVLink; EPI_PUSH in common.cuh)
//     signal_kernel    epoch += 1, st.release.sys into the neighbours' flag slots
//     interior rows    the bulk of the shard; overlaps the neighbours' waits
//   One stream, no collective call, no host round trip; two consecutive steps are captured in a
//   CUDA graph.  "Boundary" = rows a peer needs UNION rows that read a halo column, so interior
//   rows never touch the halo and a neighbour that runs one step ahead cannot overwrite data
//   that is still being read (round-1 advisor finding: with boundary = "rows peers need" only,
//   a structurally non-symmetric banded matrix raced).
//   A boundary segment that feeds MORE than two peers and covers the whole shard (a general
//   matrix: every peer needs every slice) runs with the plain epilogue on the shard's best route,
//   and its finished row blocks leave as peer copies on the copy engines, one stream per peer,
//   while later blocks still compute (csr_run_blocks); the signal follows the last copy.
//   Measured on C3 at 8 GPUs: epilogue stores to 7 peers 1.28 ms (14 M scattered 8-byte NVLink
//   writes per GPU and step), NCCL all-gather 0.83 ms, block copies 0.66 ms.
// Mode NCCL (SPMV_B200_PUSH_ALL=0, or plans with more than 8 targets per segment): boundary rows,
//   then the exchange on a second stream (ncclAllGather when every rank needs every slice and the
//   slices are equal, grouped ncclSend/ncclRecv otherwise) while the interior rows run; libnccl
//   is loaded with dlopen, not linked.
//
// Two deployments share this code: one process per GPU (peers mapped through CUDA IPC; the
// caller moves the 512-byte connection blobs between the processes, e.g. with torch.distributed
// or MPI) and one process driving all GPUs (spmv_b200_dist_group_*, peer access).
#include "internal.cuh"

#include <dlfcn.h>
#include <nccl.h> // types and prototypes only; the library is resolved at run time
#include <unistd.h>

using namespace b200;

// ======================================================================= planning
extern "C" int spmv_b200_partition_rows(int64_t M, const void *irp, int irp_bytes, int parts,
                                        int align, int64_t *cuts) {
      if (M < 0 || !irp || parts < 1 || !cuts || (irp_bytes != 4 && irp_bytes != 8))
            return fail(-EINVAL, "partition_rows: bad arguments");
      if (align < 1)
            align = 1;
      auto at = [&](int64_t r) -> int64_t {
            return irp_bytes == 4 ? (int64_t) static_cast<const int *>(irp)[r]
                                  : (int64_t) static_cast<const long long *>(irp)[r];
      };
      const double target = (double)(at(M) - at(0)) / parts;
      cuts[0] = 0;
      int64_t from = 0;
      for (int p = 1; p < parts; ++p) {
            // first r with nnz(rows[from, r)) >= target: the reference closes a part at the row
            // whose running count reaches the target and restarts the count at zero
            int64_t lo = from, hi = M;
            const double want = (double)at(from) + target;
            while (lo < hi) {
                  const int64_t mid = lo + (hi - lo) / 2;
                  if ((double)at(mid) >= want)
                        hi = mid;
                  else
                        lo = mid + 1;
            }
            int64_t r = lo;
            if (from < M)
                  r = std::max(r, from + 1);
            r = std::min<int64_t>(M, (r + align - 1) / align * align);
            cuts[p] = std::max(r, cuts[p - 1]);
            from = cuts[p];
      }
      cuts[parts] = M;
      return 0;
}

extern "C" int spmv_b200_shard_scan(int64_t r0, int64_t r1, const void *irp, int irp_bytes,
                                    const int *JA, spmv_b200_shard_desc *out) {
      if (r1 < r0 || !irp || !out || (irp_bytes != 4 && irp_bytes != 8))
            return fail(-EINVAL, "shard_scan: bad arguments");
      const int64_t M = r1 - r0;
      auto at = [&](int64_t r) -> int64_t {
            return irp_bytes == 4 ? (int64_t) static_cast<const int *>(irp)[r]
                                  : (int64_t) static_cast<const long long *>(irp)[r];
      };
      const int64_t base = at(0);
      long long cmin = r0, cmax = r1 - 1;
      long long read_lo = 0, read_hi = M; // one past the last row reading a column < r0; first row
                                          // reading a column >= r1
#pragma omp parallel for schedule(static) reduction(min : cmin, read_hi) reduction(max : cmax, read_lo)
      for (int64_t r = 0; r < M; ++r) {
            for (int64_t k = at(r) - base; k < at(r + 1) - base; ++k) {
                  const long long c = JA[k];
                  cmin = std::min(cmin, c);
                  cmax = std::max(cmax, c);
                  if (c < r0)
                        read_lo = std::max<long long>(read_lo, r + 1);
                  if (c >= r1)
                        read_hi = std::min<long long>(read_hi, r);
            }
      }
      out->r0 = r0, out->r1 = r1;
      out->c0 = M ? cmin : r0;
      out->c1 = M ? cmax + 1 : r1;
      out->read_lo = read_lo, out->read_hi = read_hi;
      return 0;
}

extern "C" int spmv_b200_stencil27_shard_desc(int nx, int ny, int nz, int z0, int z1,
                                              spmv_b200_shard_desc *out) {
      if (!out || nx < 1 || ny < 1 || nz < 1 || z0 < 0 || z1 > nz || z0 > z1)
            return fail(-EINVAL, "bad stencil geometry");
      const int64_t plane = (int64_t)nx * ny;
      out->r0 = z0 * plane, out->r1 = z1 * plane;
      out->c0 = std::max(0, z0 - 1) * plane, out->c1 = std::min(nz, z1 + 1) * plane;
      const int64_t M = out->r1 - out->r0;
      out->read_lo = z0 > 0 ? std::min(M, plane) : 0;
      out->read_hi = z1 < nz ? std::max<int64_t>(0, M - plane) : M;
      return 0;
}

namespace {

// send / recv lists and the boundary cut of one rank; mode-independent
void plan_one(int rank, int world, const spmv_b200_shard_desc *t, spmv_b200_dist_plan *p) {
      memset(p, 0, sizeof *p);
      p->rank = rank, p->world = world;
      const spmv_b200_shard_desc &me = t[rank];
      const int64_t M = me.r1 - me.r0;
      int64_t covered = M;
      for (int q = 0; q < world; ++q) {
            if (q == rank)
                  continue;
            int64_t g0 = std::max(me.c0, t[q].r0), g1 = std::min(me.c1, t[q].r1);
            if (g0 < g1) {
                  p->recv[p->n_recv++] = {q, g0, g1};
                  covered += g1 - g0;
                  p->halo_bytes += 8 * (g1 - g0);
            }
            g0 = std::max(t[q].c0, me.r0), g1 = std::min(t[q].c1, me.r1);
            if (g0 < g1)
                  p->send[p->n_send++] = {q, g0, g1};
      }
      p->covered = covered == me.c1 - me.c0;
      // local rows [0, lo) and [hi, M) are boundary: a peer needs them, or they read the halo
      int64_t lo = me.read_lo, hi = me.read_hi;
      bool inner = false;
      for (int i = 0; i < p->n_send; ++i) {
            const auto &s = p->send[i];
            if (s.g0 == me.r0)
                  lo = std::max(lo, s.g1 - me.r0);
            else if (s.g1 == me.r1)
                  hi = std::min(hi, s.g0 - me.r0);
            else
                  inner = true; // a peer needs rows from the middle
      }
      if (inner || lo >= hi)
            lo = hi = M; // everything is boundary: nothing to overlap
      p->boundary_lo = lo, p->boundary_hi = hi;
      p->n_cuts = 0;
      if (lo > 0 && lo < M)
            p->cuts[p->n_cuts++] = lo;
      if (hi > 0 && hi < M && hi != lo)
            p->cuts[p->n_cuts++] = hi;
}

struct Seg {
      int64_t r0, r1;
      bool boundary;
};
std::vector<Seg> plan_segments(const spmv_b200_dist_plan &p, int64_t M) {
      std::vector<Seg> s;
      if (M <= 0)
            return s;
      if (p.boundary_lo >= M) {
            s.push_back({0, M, true});
            return s;
      }
      if (p.boundary_lo > 0)
            s.push_back({0, p.boundary_lo, true});
      if (p.boundary_hi < M)
            s.push_back({p.boundary_hi, M, true});
      s.push_back({p.boundary_lo, p.boundary_hi, false});
      return s;
}

int max_push_targets(const spmv_b200_dist_plan &p, const spmv_b200_shard_desc &me) {
      int worst = 0;
      for (const Seg &sg : plan_segments(p, me.r1 - me.r0)) {
            if (!sg.boundary)
                  continue;
            int n = 0;
            for (int i = 0; i < p.n_send; ++i)
                  n += std::max(p.send[i].g0 - me.r0, sg.r0) < std::min(p.send[i].g1 - me.r0, sg.r1);
            worst = std::max(worst, n);
      }
      return worst;
}

} // namespace

extern "C" int spmv_b200_dist_make_plan(int rank, int world, const spmv_b200_shard_desc *table,
                                        int want_mode, spmv_b200_dist_plan *out) {
      if (!table || !out || world < 1 || world > SPMV_B200_MAX_RANKS || rank < 0 || rank >= world)
            return fail(-EINVAL, "dist_make_plan: bad arguments (at most %d ranks)",
                        SPMV_B200_MAX_RANKS);
      // the mode must be the same on every rank: decide it from everybody's plan
      bool all_gather = world > 1, equal = true, covered = true;
      int targets = 0; // most peers any boundary segment stores to
      for (int q = 0; q < world; ++q) {
            spmv_b200_dist_plan pq;
            plan_one(q, world, table, &pq);
            targets = std::max(targets, max_push_targets(pq, table[q]));
            covered = covered && pq.covered;
            all_gather = all_gather && table[q].c0 == table[0].r0 && table[q].c1 == table[world - 1].r1;
            equal = equal && table[q].r1 - table[q].r0 == table[0].r1 - table[0].r0 &&
                    table[q].r0 == table[0].r0 + q * (table[0].r1 - table[0].r0);
      }
      if (!covered)
            return fail(-EINVAL, "dist_make_plan: row ranges do not tile the needed column ranges");
      plan_one(rank, world, table, out);
      out->all_gather = all_gather && equal;
      if (want_mode == SPMV_B200_DIST_PUSH && targets > kMaxPush)
            return fail(-EINVAL, "dist_make_plan: this partition needs more than %d push targets "
                                 "per boundary segment; use SPMV_B200_DIST_NCCL or _AUTO", kMaxPush);
      // AUTO: halo plans (at most two peers per boundary segment) push.  Plans in which a segment
      // feeds more peers -- a general matrix: every peer needs every slice -- push as well (the
      // all-gather fused into the SpMV epilogue, stores to up to 7 peers over NVLink) unless
      // SPMV_B200_PUSH_ALL=0 sends them through the NCCL exchange.
      bool push = targets <= 2;
      if (!push && targets <= kMaxPush) {
            const char *env = getenv("SPMV_B200_PUSH_ALL");
            push = !(env && !strcmp(env, "0"));
      }
      out->mode = want_mode == SPMV_B200_DIST_AUTO ? (push ? SPMV_B200_DIST_PUSH : SPMV_B200_DIST_NCCL) : want_mode;
      return 0;
}

// ================================================= cross-GPU step ordering + IPC
// Every rank owns `flags[world]` and an `epoch` word in its own HBM, mapped into the neighbours.
// After the boundary-row kernel of a step has pushed its halo rows into the neighbours,
// signal_kernel bumps the local epoch and stores it into slot [my_rank] of every neighbour's
// flags; before the next step's boundary rows, wait_kernel spins (bounded) until every
// neighbour's slot has reached the local epoch.
namespace {

struct PeerSlots {
      int n;
      unsigned long long *slot[SPMV_B200_MAX_RANKS];
};

__global__ void signal_kernel(unsigned long long *epoch, PeerSlots peers) {
      if (threadIdx.x != 0 || blockIdx.x != 0)
            return;
      const unsigned long long e = *epoch + 1;
      *epoch = e;
      __threadfence_system(); // halo rows pushed by earlier kernels are visible first
      for (int i = 0; i < peers.n; ++i)
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peers.slot[i]), "l"(e)
                         : "memory");
}

__global__ void wait_kernel(const unsigned long long *epoch, PeerSlots mine,
                            unsigned long long max_spins, int *error) {
      if (threadIdx.x != 0 || blockIdx.x != 0)
            return;
      if (*(volatile int *)error != 0)
            return; // a previous wait already gave up: do not stall every later step
      const unsigned long long want = *epoch;
      for (int i = 0; i < mine.n; ++i) {
            unsigned long long spins = 0, v;
            for (;;) {
                  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine.slot[i])
                               : "memory");
                  if (v >= want)
                        break;
                  if (++spins > max_spins) { // never hang the GPU: report and go on
                        atomicExch(error, 1 + i);
                        return;
                  }
                  // plain polling: __nanosleep() rounds up to a scheduler quantum that is
                  // longer than a whole SpMV step of a 128^3 slab
            }
      }
}

} // namespace

extern "C" int spmv_b200_signal_peers(void *d_epoch, int n, void *const *d_peer_slots,
                                      void *stream) {
      if (n < 0 || n > SPMV_B200_MAX_RANKS)
            return fail(-EINVAL, "signal_peers: too many peers");
      PeerSlots p{};
      p.n = n;
      for (int i = 0; i < n; ++i)
            p.slot[i] = static_cast<unsigned long long *>(d_peer_slots[i]);
      signal_kernel<<<1, 32, 0, as_stream(stream)>>>(static_cast<unsigned long long *>(d_epoch), p);
      ++g_counters.launches;
      B200_CUDA(cudaGetLastError());
      return 0;
}

extern "C" int spmv_b200_wait_peers(const void *d_epoch, int n, void *const *d_my_slots,
                                    uint64_t max_spins, int *d_error, void *stream) {
      if (n < 0 || n > SPMV_B200_MAX_RANKS)
            return fail(-EINVAL, "wait_peers: too many peers");
      PeerSlots p{};
      p.n = n;
      for (int i = 0; i < n; ++i)
            p.slot[i] = static_cast<unsigned long long *>(d_my_slots[i]);
      wait_kernel<<<1, 32, 0, as_stream(stream)>>>(static_cast<const unsigned long long *>(d_epoch),
                                                   p, max_spins, d_error);
      ++g_counters.launches;
      B200_CUDA(cudaGetLastError());
      return 0;
}

extern "C" int spmv_b200_ipc_export(void *d_ptr, unsigned char *handle64) {
      static_assert(sizeof(cudaIpcMemHandle_t) == SPMV_B200_IPC_HANDLE_BYTES, "handle size");
      cudaIpcMemHandle_t hd;
      B200_CUDA(cudaIpcGetMemHandle(&hd, d_ptr));
      memcpy(handle64, &hd, sizeof hd);
      return 0;
}
extern "C" int spmv_b200_ipc_open(const unsigned char *handle64, void **d_ptr_out) {
      cudaIpcMemHandle_t hd;
      memcpy(&hd, handle64, sizeof hd);
      B200_CUDA(cudaIpcOpenMemHandle(d_ptr_out, hd, cudaIpcMemLazyEnablePeerAccess));
      return 0;
}
extern "C" int spmv_b200_ipc_close(void *d_ptr) {
      B200_CUDA(cudaIpcCloseMemHandle(d_ptr));
      return 0;
}
extern "C" int spmv_b200_enable_peer(int peer_device) {
      cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) {
            cudaGetLastError();
            return 0;
      }
      B200_CUDA(e);
      return 0;
}

// ========================================================================= NCCL
namespace {

struct NcclApi {
      void *lib = nullptr;
      decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
      decltype(&ncclCommInitRank) CommInitRank = nullptr;
      decltype(&ncclCommDestroy) CommDestroy = nullptr;
      decltype(&ncclAllGather) AllGather = nullptr;
      decltype(&ncclSend) Send = nullptr;
      decltype(&ncclRecv) Recv = nullptr;
      decltype(&ncclGroupStart) GroupStart = nullptr;
      decltype(&ncclGroupEnd) GroupEnd = nullptr;
      decltype(&ncclGetErrorString) GetErrorString = nullptr;
} g_nccl;

int nccl_load() {
      static std::once_flag once;
      static int rc = 0;
      std::call_once(once, [] {
            const char *names[] = {"libnccl.so.2", "libnccl.so"};
            for (const char *n : names)
                  if ((g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL)))
                        break;
            if (!g_nccl.lib) {
                  rc = fail(-ENOENT, "libnccl.so.2 not found (%s): the NCCL exchange is unavailable",
                            dlerror());
                  return;
            }
#define SYM(field, name)                                                                           \
      g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.lib, name));            \
      if (!g_nccl.field)                                                                           \
            rc = fail(-ENOENT, "libnccl: missing symbol %s", name);
            SYM(GetUniqueId, "ncclGetUniqueId")
            SYM(CommInitRank, "ncclCommInitRank")
            SYM(CommDestroy, "ncclCommDestroy")
            SYM(AllGather, "ncclAllGather")
            SYM(Send, "ncclSend")
            SYM(Recv, "ncclRecv")
            SYM(GroupStart, "ncclGroupStart")
            SYM(GroupEnd, "ncclGroupEnd")
            SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
      });
      return rc;
}

#define B200_NCCL(call)                                                                            \
      do {                                                                                         \
            ncclResult_t r_ = (call);                                                              \
            if (r_ != ncclSuccess)                                                                 \
                  return ::b200::fail(-EIO, "%s failed: %s", #call, g_nccl.GetErrorString(r_));    \
      } while (0)

// layout of a connection blob (SPMV_B200_DIST_BLOB_BYTES)
struct Blob {
      uint32_t magic;
      int32_t rank, device, pid;
      uint64_t raw_arena, x_bytes; // same-process peers use the pointer as it is
      cudaIpcMemHandle_t ipc_arena; // other processes map the arena through CUDA IPC
      int32_t has_nccl_id;
      ncclUniqueId nccl_id; // from rank 0, NCCL mode
};
static_assert(sizeof(Blob) <= SPMV_B200_DIST_BLOB_BYTES, "blob size");
constexpr uint32_t kBlobMagic = 0xB2005D15u;

struct Peer {
      double *x[2] = {nullptr, nullptr};
      unsigned long long *flags = nullptr;
      void *ipc_base = nullptr; // non-null: mapped with cudaIpcOpenMemHandle
      int64_t c0 = 0;
};
constexpr size_t kCtlBytes = 4096;
// all-gather plans: the step's rows become final in up to this many blocks, each handed to the
// copy engines while the later ones still compute
constexpr int kGatherBlocks = 4;

} // namespace

struct spmv_b200_dist {
      spmv_b200_dist_plan plan;
      spmv_b200_shard_desc table[SPMV_B200_MAX_RANKS];
      spmv_b200_shard_desc me;
      int rank = 0, world = 1, device = 0, kernel = 4, wpb = 4, mode = SPMV_B200_DIST_PUSH;
      int64_t M = 0, n_local = 0, own0 = 0;
      spmv_b200_csr *shard = nullptr;
      bool own_shard = false;
      // ONE device allocation per rank (one IPC handle): control block | X[0] | X[1]
      unsigned char *arena = nullptr;
      size_t x_bytes = 0;
      double *X[2] = {nullptr, nullptr};
      // control block: flags[MAX_RANKS] | epoch | error
      unsigned long long *ctl = nullptr;
      unsigned long long *flags() { return ctl; }
      unsigned long long *epoch() { return ctl + SPMV_B200_MAX_RANKS; }
      int *err() { return reinterpret_cast<int *>(ctl + SPMV_B200_MAX_RANKS + 1); }
      Peer peers[SPMV_B200_MAX_RANKS];
      std::vector<int> neighbours;
      PeerSlots peer_slots{}, my_slots{};
      std::vector<Seg> segs;
      cudaStream_t st = nullptr, comm = nullptr;
      cudaEvent_t ev_b = nullptr, ev_done = nullptr;
      cudaEvent_t ev_blk[kGatherBlocks] = {nullptr}; // all-gather by copy engines: one per finished row block
      cudaStream_t peer_st[kMaxPush] = {nullptr};    // ... one stream per target, so that the copies to
      cudaEvent_t peer_done[kMaxPush] = {nullptr};   //     different peers run on different copy engines
      cudaGraphExec_t graph = nullptr;
      bool graph_failed = false, connected = false;
      int eager_steps = 0;
      long long step_no = 0;
      ncclComm_t nccl = nullptr;
      ncclUniqueId nccl_id{};
      bool in_group = false; // single-process group: NCCL calls are bracketed by the group code
      std::vector<cudaEvent_t> t_ev;
};

namespace {

double *own(spmv_b200_dist *d, int buf) { return d->X[buf] + d->own0; }

// row blocks per step of an all-gather plan (SPMV_B200_GATHER_BLOCKS = 1 .. kGatherBlocks)
int gather_blocks() {
      static int n = 0;
      if (!n) {
            const char *env = getenv("SPMV_B200_GATHER_BLOCKS");
            n = env && atoi(env) > 0 ? std::min(atoi(env), kGatherBlocks) : 2; // C3 at 8 GPUs: 1 / 2 / 4 blocks -> 0.744 / 0.659 / 0.696 ms
      }
      return n;
}

int nccl_exchange(spmv_b200_dist *d, int buf, cudaStream_t st) {
      const spmv_b200_dist_plan &P = d->plan;
      if (d->world == 1)
            return 0;
      if (P.all_gather) {
            B200_NCCL(g_nccl.AllGather(own(d, buf), d->X[buf], (size_t)d->M, ncclFloat64, d->nccl, st));
            return 0;
      }
      if (!d->in_group)
            B200_NCCL(g_nccl.GroupStart());
      for (int i = 0; i < P.n_send; ++i)
            B200_NCCL(g_nccl.Send(d->X[buf] + (P.send[i].g0 - d->me.c0),
                                  (size_t)(P.send[i].g1 - P.send[i].g0), ncclFloat64, P.send[i].peer,
                                  d->nccl, st));
      for (int i = 0; i < P.n_recv; ++i)
            B200_NCCL(g_nccl.Recv(d->X[buf] + (P.recv[i].g0 - d->me.c0),
                                  (size_t)(P.recv[i].g1 - P.recv[i].g0), ncclFloat64, P.recv[i].peer,
                                  d->nccl, st));
      if (!d->in_group)
            B200_NCCL(g_nccl.GroupEnd());
      return 0;
}

// rows of a boundary segment that peers need -> where they live in the peers' x buffers
int push_args_for(spmv_b200_dist *d, const Seg &sg, int dst, EpiArgs *out) {
      *out = EpiArgs{};
      const spmv_b200_dist_plan &P = d->plan;
      for (int i = 0; i < P.n_send; ++i) {
            const int64_t a = std::max(P.send[i].g0 - d->me.r0, sg.r0),
                          b = std::min(P.send[i].g1 - d->me.r0, sg.r1);
            if (a >= b)
                  continue;
            if (out->n_push >= kMaxPush)
                  return fail(-EINVAL, "push epilogue supports %d peers per boundary segment", kMaxPush);
            const Peer &pr = d->peers[P.send[i].peer];
            out->row0[out->n_push] = a;
            out->row1[out->n_push] = b;
            // local row a is global index r0 + a: its place in the peer's buffer
            out->dst[out->n_push] = pr.x[dst] + (d->me.r0 + a - pr.c0);
            ++out->n_push;
      }
      return 0;
}

// One iteration, queued on d->st (and d->comm in NCCL mode).  phase: 0 = everything;
// 1 = up to and including the boundary rows, 2 = the exchange, 3 = the rest (the
// single-process group brackets phase 2 of all ranks with ncclGroupStart/End).
int dist_step(spmv_b200_dist *d, int phase = 0) {
      const int src = (int)(d->step_no % 2), dst = 1 - src;
      const double *x = d->X[src];
      double *y = own(d, dst);
      cudaStream_t st = d->st;
      if (d->mode == SPMV_B200_DIST_PUSH) {
            if (d->world > 1) {
                  wait_kernel<<<1, 32, 0, st>>>(d->epoch(), d->my_slots, 1ull << 24, d->err());
                  ++g_counters.launches;
            }
            for (const Seg &sg : d->segs) {
                  if (!sg.boundary)
                        continue;
                  EpiArgs e;
                  int rc = push_args_for(d, sg, dst, &e);
                  if (!rc && e.n_push > 2 && sg.r0 == 0 && sg.r1 == d->M) {
                        // A general matrix: every peer needs this rank's whole next slice.  Stores
                        // from the epilogue would leave as 8-byte writes in the row order of the
                        // sorted slices / row lists (measured at 8 GPUs on C3: 14 M remote writes
                        // per GPU and step, 1.28 ms against 0.83 ms through NCCL), so the finished
                        // row blocks go out as peer copies on the copy engines instead, block b
                        // travelling while block b+1 computes; the shard keeps its best
                        // single-GPU route (column panels, virtual rows).
                        int nb = 0;
                        cudaError_t ce = cudaSuccess;
                        for (int i = 0; i < e.n_push && ce == cudaSuccess; ++i)
                              if (!d->peer_st[i]) {
                                    int lo_p = 0, hi_p = 0;
                                    cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
                                    ce = cudaStreamCreateWithPriority(&d->peer_st[i], cudaStreamNonBlocking, hi_p);
                                    if (ce == cudaSuccess)
                                          ce = cudaEventCreateWithFlags(&d->peer_done[i], cudaEventDisableTiming);
                              }
                        auto done = [&](long long a, long long b) {
                              if (nb >= kGatherBlocks || a >= b)
                                    return;
                              cudaEvent_t ev = d->ev_blk[nb++];
                              if (ce == cudaSuccess)
                                    ce = cudaEventRecord(ev, st);
                              for (int i = 0; i < e.n_push && ce == cudaSuccess; ++i) {
                                    const long long lo = std::max<long long>(a, e.row0[i]),
                                                    hi = std::min<long long>(b, e.row1[i]);
                                    if (lo >= hi)
                                          continue;
                                    ce = cudaStreamWaitEvent(d->peer_st[i], ev, 0);
                                    if (ce == cudaSuccess)
                                          ce = cudaMemcpyAsync(e.dst[i] + (lo - e.row0[i]), y + lo,
                                                               (size_t)(hi - lo) * sizeof(double),
                                                               cudaMemcpyDeviceToDevice, d->peer_st[i]);
                              }
                        };
                        rc = ce == cudaSuccess ? csr_run_blocks(d->shard, d->kernel, d->wpb, x, y, st, gather_blocks(), done)
                                               : -EIO;
                        if (ce != cudaSuccess)
                              rc = fail(-EIO, "all-gather copies failed: %s", cudaGetErrorString(ce));
                        for (int i = 0; i < e.n_push && !rc && nb > 0; ++i) { // join: the signal follows the last copy
                              B200_CUDA(cudaEventRecord(d->peer_done[i], d->peer_st[i]));
                              B200_CUDA(cudaStreamWaitEvent(st, d->peer_done[i], 0));
                        }
                        if (rc)
                              return rc;
                        continue;
                  }
                  rc = rc ? rc : csr_run(d->shard, d->kernel, d->wpb, sg.r0, sg.r1, x, y,
                                         e.n_push ? EPI_PUSH : EPI_PLAIN, e, st);
                  if (rc)
                        return rc;
            }
            if (d->world > 1) {
                  signal_kernel<<<1, 32, 0, st>>>(d->epoch(), d->peer_slots);
                  ++g_counters.launches;
            }
            for (const Seg &sg : d->segs)
                  if (!sg.boundary) {
                        int rc = csr_run(d->shard, d->kernel, d->wpb, sg.r0, sg.r1, x, y, EPI_PLAIN,
                                         EpiArgs{}, st);
                        if (rc)
                              return rc;
                  }
            ++d->step_no;
            return 0;
      }
      // NCCL mode
      bool has_interior = false;
      for (const Seg &sg : d->segs)
            has_interior = has_interior || !sg.boundary;
      if (phase == 0 || phase == 1) {
            for (const Seg &sg : d->segs)
                  if (sg.boundary) {
                        int rc = csr_run(d->shard, d->kernel, d->wpb, sg.r0, sg.r1, x, y, EPI_PLAIN,
                                         EpiArgs{}, st);
                        if (rc)
                              return rc;
                  }
            if (has_interior) {
                  B200_CUDA(cudaEventRecord(d->ev_b, st));
                  B200_CUDA(cudaStreamWaitEvent(d->comm, d->ev_b, 0));
            }
      }
      if (phase == 0 || phase == 2) {
            int rc = nccl_exchange(d, dst, has_interior ? d->comm : st);
            if (rc)
                  return rc;
      }
      if (phase == 0 || phase == 3) {
            if (has_interior) {
                  B200_CUDA(cudaEventRecord(d->ev_done, d->comm));
                  for (const Seg &sg : d->segs)
                        if (!sg.boundary) {
                              int rc = csr_run(d->shard, d->kernel, d->wpb, sg.r0, sg.r1, x, y,
                                               EPI_PLAIN, EpiArgs{}, st);
                              if (rc)
                                    return rc;
                        }
                  B200_CUDA(cudaStreamWaitEvent(st, d->ev_done, 0));
            }
            ++d->step_no;
      }
      return 0;
}

// Capture two consecutive steps (one per x buffer) once the launch plans exist.
void try_capture(spmv_b200_dist *d) {
      if (d->graph || d->graph_failed || d->mode != SPMV_B200_DIST_PUSH || d->step_no % 2 != 0 ||
          d->eager_steps < 2)
            return;
      const char *env = getenv("SPMV_B200_GRAPH");
      if (env && !strcmp(env, "0")) {
            d->graph_failed = true;
            return;
      }
      const long long saved = d->step_no, saved_launches = g_counters.launches;
      cudaGraph_t g = nullptr;
      bool ok = cudaStreamBeginCapture(d->st, cudaStreamCaptureModeRelaxed) == cudaSuccess;
      if (ok) {
            ok = dist_step(d) == 0 && dist_step(d) == 0;
            ok = (cudaStreamEndCapture(d->st, &g) == cudaSuccess) && ok && g;
      }
      if (ok)
            ok = cudaGraphInstantiate(&d->graph, g, 0) == cudaSuccess;
      if (g)
            cudaGraphDestroy(g);
      d->step_no = saved;
      g_counters.launches = saved_launches; // captured, not run
      if (!ok) {
            cudaGetLastError();
            d->graph = nullptr;
            d->graph_failed = true;
      }
}

int dist_iterate(spmv_b200_dist *d, int k) {
      B200_CUDA(cudaSetDevice(d->device));
      while (k > 0) {
            if (k >= 2)
                  try_capture(d);
            if (d->graph && k >= 2 && d->step_no % 2 == 0) {
                  B200_CUDA(cudaGraphLaunch(d->graph, d->st));
                  g_counters.launches += 2 * ((long long)d->segs.size() + (d->world > 1 ? 2 : 0));
                  d->step_no += 2;
                  k -= 2;
            } else {
                  int rc = dist_step(d);
                  if (rc)
                        return rc;
                  ++d->eager_steps;
                  --k;
            }
      }
      return 0;
}

} // namespace

// ================================================================== per-rank API
extern "C" spmv_b200_dist *spmv_b200_dist_create(const spmv_b200_dist_plan *plan,
                                                 const spmv_b200_shard_desc *table,
                                                 spmv_b200_csr *shard, int kernel, int wpb) {
      if (ensure_device())
            return nullptr;
      if (!plan || !table || !shard || plan->world < 1 || plan->world > SPMV_B200_MAX_RANKS) {
            fail(-EINVAL, "dist_create: bad arguments");
            return nullptr;
      }
      const spmv_b200_shard_desc &me = table[plan->rank];
      if (shard->M != me.r1 - me.r0 || shard->N != me.c1 - me.c0 || shard->col_offset != me.c0) {
            fail(-EINVAL, "dist_create: the shard (M=%lld N=%lld col_offset=%lld) does not match "
                          "its descriptor (rows %lld, cols [%lld,%lld))",
                 shard->M, shard->N, shard->col_offset, (long long)(me.r1 - me.r0),
                 (long long)me.c0, (long long)me.c1);
            return nullptr;
      }
      if (plan->mode == SPMV_B200_DIST_NCCL && plan->world > 1 && nccl_load())
            return nullptr;
      auto *d = new spmv_b200_dist();
      d->plan = *plan;
      memcpy(d->table, table, sizeof(spmv_b200_shard_desc) * (size_t)plan->world);
      d->me = me;
      d->rank = plan->rank, d->world = plan->world, d->mode = plan->mode;
      d->kernel = kernel, d->wpb = clamp_wpb(wpb);
      d->M = me.r1 - me.r0, d->n_local = me.c1 - me.c0, d->own0 = me.r0 - me.c0;
      d->shard = shard;
      d->segs = plan_segments(*plan, d->M);
      cudaGetDevice(&d->device);
      // at least 2 MiB so the arena never shares a driver block with unrelated small allocations
      d->x_bytes = (((size_t)d->n_local + 32) * sizeof(double) + 255) & ~(size_t)255;
      const size_t arena_bytes = std::max<size_t>(kCtlBytes + 2 * d->x_bytes, 2u << 20);
      bool ok = cudaMalloc(&d->arena, arena_bytes) == cudaSuccess &&
                cudaMemset(d->arena, 0, arena_bytes) == cudaSuccess;
      if (ok) {
            d->ctl = reinterpret_cast<unsigned long long *>(d->arena);
            d->X[0] = reinterpret_cast<double *>(d->arena + kCtlBytes);
            d->X[1] = reinterpret_cast<double *>(d->arena + kCtlBytes + d->x_bytes);
      }
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);
      ok = ok && cudaStreamCreateWithFlags(&d->st, cudaStreamNonBlocking) == cudaSuccess &&
           // high priority: when the exchange and the (SM-filling, persistent) interior kernel
           // become runnable together, the exchange gets its SM first
           cudaStreamCreateWithPriority(&d->comm, cudaStreamNonBlocking, hi) == cudaSuccess &&
           cudaEventCreateWithFlags(&d->ev_b, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&d->ev_done, cudaEventDisableTiming) == cudaSuccess;
      for (int i = 0; i < kGatherBlocks; ++i)
            ok = ok && cudaEventCreateWithFlags(&d->ev_blk[i], cudaEventDisableTiming) == cudaSuccess;
      if (!ok) {
            fail(-ENOMEM, "dist_create: %s", cudaGetErrorString(cudaGetLastError()));
            spmv_b200_dist_destroy(d);
            return nullptr;
      }
      for (int i = 0; i < plan->n_send; ++i)
            d->neighbours.push_back(plan->send[i].peer);
      for (int i = 0; i < plan->n_recv; ++i)
            d->neighbours.push_back(plan->recv[i].peer);
      std::sort(d->neighbours.begin(), d->neighbours.end());
      d->neighbours.erase(std::unique(d->neighbours.begin(), d->neighbours.end()),
                          d->neighbours.end());
      if (d->world == 1)
            d->connected = true;
      return d;
}

extern "C" int spmv_b200_dist_export(spmv_b200_dist *d, unsigned char *blob) {
      if (!d || !blob)
            return fail(-EINVAL, "dist_export: null argument");
      B200_CUDA(cudaSetDevice(d->device));
      Blob b{};
      b.magic = kBlobMagic;
      b.rank = d->rank, b.device = d->device, b.pid = (int32_t)getpid();
      b.raw_arena = (uint64_t)(uintptr_t)d->arena;
      b.x_bytes = d->x_bytes;
      B200_CUDA(cudaIpcGetMemHandle(&b.ipc_arena, d->arena));
      if (d->mode == SPMV_B200_DIST_NCCL && d->rank == 0 && d->world > 1) {
            B200_NCCL(g_nccl.GetUniqueId(&b.nccl_id));
            b.has_nccl_id = 1;
      }
      memset(blob, 0, SPMV_B200_DIST_BLOB_BYTES);
      memcpy(blob, &b, sizeof b);
      return 0;
}

extern "C" int spmv_b200_dist_connect(spmv_b200_dist *d, const unsigned char *blobs) {
      if (!d || !blobs)
            return fail(-EINVAL, "dist_connect: null argument");
      B200_CUDA(cudaSetDevice(d->device));
      const int32_t my_pid = (int32_t)getpid();
      for (int p : d->neighbours) {
            Blob b;
            memcpy(&b, blobs + (size_t)p * SPMV_B200_DIST_BLOB_BYTES, sizeof b);
            if (b.magic != kBlobMagic || b.rank != p)
                  return fail(-EINVAL, "dist_connect: blob %d is not from rank %d", p, p);
            Peer &pr = d->peers[p];
            pr.c0 = d->table[p].c0;
            if (d->mode != SPMV_B200_DIST_PUSH)
                  continue;
            if (b.pid == my_pid) { // same process: peer access, pointers as they are
                  if (b.device != d->device) {
                        int can = 0;
                        cudaDeviceCanAccessPeer(&can, d->device, b.device);
                        if (!can)
                              return fail(-ENOTSUP, "GPU %d cannot access GPU %d", d->device, b.device);
                        int rc = spmv_b200_enable_peer(b.device);
                        if (rc)
                              return rc;
                  }
            }
            unsigned char *base = (unsigned char *)(uintptr_t)b.raw_arena;
            if (b.pid != my_pid) {
                  void *q = nullptr;
                  B200_CUDA(cudaIpcOpenMemHandle(&q, b.ipc_arena, cudaIpcMemLazyEnablePeerAccess));
                  pr.ipc_base = q;
                  base = (unsigned char *)q;
            }
            pr.flags = (unsigned long long *)base;
            pr.x[0] = (double *)(base + kCtlBytes);
            pr.x[1] = (double *)(base + kCtlBytes + b.x_bytes);
      }
      d->peer_slots.n = d->my_slots.n = 0;
      if (d->mode == SPMV_B200_DIST_PUSH)
            for (int p : d->neighbours) {
                  // slot [rank] of each neighbour's flags (I write), slot [p] of mine (p writes)
                  d->peer_slots.slot[d->peer_slots.n++] = d->peers[p].flags + d->rank;
                  d->my_slots.slot[d->my_slots.n++] = d->flags() + p;
            }
      if (d->mode == SPMV_B200_DIST_NCCL && d->world > 1) {
            Blob b0;
            memcpy(&b0, blobs, sizeof b0);
            if (!b0.has_nccl_id)
                  return fail(-EINVAL, "dist_connect: rank 0's blob carries no NCCL id");
            B200_NCCL(g_nccl.CommInitRank(&d->nccl, d->world, b0.nccl_id, d->rank));
      }
      d->connected = true;
      return 0;
}

// x_0: the caller's own slice (host or device memory).  Every rank calls this collectively while
// the whole job is idle (synchronise + barrier first): it resets the step counter and sends the
// boundary values to the neighbours (peer copies + an epoch signal in PUSH mode, the NCCL
// exchange otherwise).
extern "C" int spmv_b200_dist_set_x(spmv_b200_dist *d, const double *x_own) {
      if (!d || !x_own || !d->connected)
            return fail(-EINVAL, "dist_set_x: not connected, or null x");
      B200_CUDA(cudaSetDevice(d->device));
      d->step_no = 0;
      B200_CUDA(cudaMemcpyAsync(own(d, 0), x_own, (size_t)d->M * 8, cudaMemcpyDefault, d->st));
      if (d->world == 1)
            return 0;
      if (d->mode == SPMV_B200_DIST_PUSH) {
            const spmv_b200_dist_plan &P = d->plan;
            for (int i = 0; i < P.n_send; ++i) {
                  const Peer &pr = d->peers[P.send[i].peer];
                  B200_CUDA(cudaMemcpyAsync(pr.x[0] + (P.send[i].g0 - pr.c0),
                                            d->X[0] + (P.send[i].g0 - d->me.c0),
                                            (size_t)(P.send[i].g1 - P.send[i].g0) * 8,
                                            cudaMemcpyDefault, d->st));
            }
            signal_kernel<<<1, 32, 0, d->st>>>(d->epoch(), d->peer_slots);
            ++g_counters.launches;
            B200_CUDA(cudaGetLastError());
            return 0;
      }
      return nccl_exchange(d, 0, d->st);
}

extern "C" int spmv_b200_dist_iterate(spmv_b200_dist *d, int k) {
      if (!d || !d->connected)
            return fail(-EINVAL, "dist_iterate: not connected");
      return dist_iterate(d, k);
}

// `reps` back-to-back regions of `k` steps each, every region bracketed by CUDA events on the
// dist's stream; returns after the stream has drained.  ms_out[reps].
extern "C" int spmv_b200_dist_time(spmv_b200_dist *d, int k, int reps, double *ms_out) {
      if (!d || !d->connected || reps < 1 || !ms_out)
            return fail(-EINVAL, "dist_time: bad arguments");
      B200_CUDA(cudaSetDevice(d->device));
      while ((int)d->t_ev.size() < 2 * reps) {
            cudaEvent_t e;
            B200_CUDA(cudaEventCreate(&e));
            d->t_ev.push_back(e);
      }
      for (int r = 0; r < reps; ++r) {
            B200_CUDA(cudaEventRecord(d->t_ev[2 * r], d->st));
            int rc = dist_iterate(d, k);
            if (rc)
                  return rc;
            B200_CUDA(cudaEventRecord(d->t_ev[2 * r + 1], d->st));
      }
      int rc = spmv_b200_dist_sync(d);
      if (rc)
            return rc;
      for (int r = 0; r < reps; ++r) {
            float ms = 0.f;
            B200_CUDA(cudaEventElapsedTime(&ms, d->t_ev[2 * r], d->t_ev[2 * r + 1]));
            ms_out[r] = ms;
      }
      return 0;
}

extern "C" int spmv_b200_dist_sync(spmv_b200_dist *d) {
      if (!d)
            return fail(-EINVAL, "null dist handle");
      B200_CUDA(cudaSetDevice(d->device));
      B200_CUDA(cudaStreamSynchronize(d->st));
      B200_CUDA(cudaStreamSynchronize(d->comm));
      int e = 0;
      B200_CUDA(cudaMemcpy(&e, d->err(), sizeof e, cudaMemcpyDeviceToHost));
      if (e)
            return fail(-ETIMEDOUT, "rank %d: wait on neighbour slot %d timed out", d->rank, e - 1);
      return 0;
}

extern "C" double *spmv_b200_dist_x(spmv_b200_dist *d) {
      return d ? own(d, (int)(d->step_no % 2)) : nullptr;
}
extern "C" double *spmv_b200_dist_xlocal(spmv_b200_dist *d) {
      return d ? d->X[d->step_no % 2] : nullptr;
}
extern "C" void *spmv_b200_dist_stream(spmv_b200_dist *d) { return d ? d->st : nullptr; }
extern "C" int64_t spmv_b200_dist_steps(const spmv_b200_dist *d) { return d ? d->step_no : -1; }
extern "C" int spmv_b200_dist_mode(const spmv_b200_dist *d) { return d ? d->mode : -1; }
extern "C" int spmv_b200_dist_has_graph(const spmv_b200_dist *d) { return d && d->graph ? 1 : 0; }

extern "C" int spmv_b200_dist_get_x(spmv_b200_dist *d, double *host_out) {
      if (!d || !host_out)
            return fail(-EINVAL, "dist_get_x: null argument");
      int rc = spmv_b200_dist_sync(d);
      if (rc)
            return rc;
      B200_CUDA(cudaMemcpy(host_out, spmv_b200_dist_x(d), (size_t)d->M * 8, cudaMemcpyDeviceToHost));
      g_counters.d2h += d->M * 8;
      return 0;
}

extern "C" void spmv_b200_dist_destroy(spmv_b200_dist *d) {
      if (!d)
            return;
      cudaSetDevice(d->device);
      if (d->st)
            cudaStreamSynchronize(d->st);
      if (d->comm)
            cudaStreamSynchronize(d->comm);
      if (d->graph)
            cudaGraphExecDestroy(d->graph);
      if (d->nccl)
            g_nccl.CommDestroy(d->nccl);
      for (auto &pr : d->peers)
            if (pr.ipc_base)
                  cudaIpcCloseMemHandle(pr.ipc_base);
      for (auto e : d->t_ev)
            cudaEventDestroy(e);
      if (d->ev_b)
            cudaEventDestroy(d->ev_b);
      if (d->ev_done)
            cudaEventDestroy(d->ev_done);
      for (cudaEvent_t e : d->ev_blk)
            if (e)
                  cudaEventDestroy(e);
      for (int i = 0; i < kMaxPush; ++i) {
            if (d->peer_st[i]) {
                  cudaStreamSynchronize(d->peer_st[i]);
                  cudaStreamDestroy(d->peer_st[i]);
            }
            if (d->peer_done[i])
                  cudaEventDestroy(d->peer_done[i]);
      }
      if (d->st)
            cudaStreamDestroy(d->st);
      if (d->comm)
            cudaStreamDestroy(d->comm);
      cudaFree(d->arena);
      if (d->own_shard)
            spmv_b200_csr_destroy(d->shard);
      cudaGetLastError();
      delete d;
}

// ======================================================= one process, all GPUs
struct spmv_b200_dist_group {
      int n = 0;
      int64_t N = 0;
      std::vector<spmv_b200_dist *> ranks;
      std::vector<spmv_b200_shard_desc> table;
      std::vector<cudaEvent_t> ev0, ev1;
      int home_device = 0;
};

namespace {

int group_wire(spmv_b200_dist_group *g) {
      std::vector<unsigned char> blobs((size_t)g->n * SPMV_B200_DIST_BLOB_BYTES);
      for (int r = 0; r < g->n; ++r) {
            int rc = spmv_b200_dist_export(g->ranks[r], blobs.data() + (size_t)r * SPMV_B200_DIST_BLOB_BYTES);
            if (rc)
                  return rc;
      }
      const bool nccl = g->ranks[0]->mode == SPMV_B200_DIST_NCCL && g->n > 1;
      if (nccl)
            B200_NCCL(g_nccl.GroupStart());
      for (int r = 0; r < g->n; ++r) {
            g->ranks[r]->in_group = true;
            int rc = spmv_b200_dist_connect(g->ranks[r], blobs.data());
            if (rc)
                  return rc;
      }
      if (nccl)
            B200_NCCL(g_nccl.GroupEnd());
      g->ev0.resize(g->n), g->ev1.resize(g->n);
      for (int r = 0; r < g->n; ++r) {
            B200_CUDA(cudaSetDevice(g->ranks[r]->device));
            B200_CUDA(cudaEventCreate(&g->ev0[r]));
            B200_CUDA(cudaEventCreate(&g->ev1[r]));
      }
      return 0;
}

spmv_b200_dist_group *group_fail(spmv_b200_dist_group *g) {
      spmv_b200_dist_group_destroy(g);
      return nullptr;
}

} // namespace

extern "C" spmv_b200_dist_group *spmv_b200_dist_group_create(const sparse_csr *A, int n_gpus,
                                                             int kernel, int wpb, int mode) {
      if (ensure_device())
            return nullptr;
      if (!A || n_gpus < 1 || n_gpus > SPMV_B200_MAX_RANKS || n_gpus > spmv_b200_device_count() ||
          A->M != A->N) {
            fail(-EINVAL, "dist_group_create: need a square matrix and 1..%d GPUs (%d visible)",
                 SPMV_B200_MAX_RANKS, spmv_b200_device_count());
            return nullptr;
      }
      auto *g = new spmv_b200_dist_group();
      cudaGetDevice(&g->home_device);
      g->n = n_gpus, g->N = A->N;
      std::vector<int64_t> cuts((size_t)n_gpus + 1);
      if (spmv_b200_partition_rows(A->M, A->IRP, 4, n_gpus, kHack, cuts.data()))
            return group_fail(g);
      g->table.resize(n_gpus);
      for (int r = 0; r < n_gpus; ++r)
            if (spmv_b200_shard_scan(cuts[r], cuts[r + 1], A->IRP + cuts[r], 4, A->JA + A->IRP[cuts[r]],
                                     &g->table[r]))
                  return group_fail(g);
      for (int r = 0; r < n_gpus; ++r) {
            spmv_b200_dist_plan plan;
            if (spmv_b200_dist_make_plan(r, n_gpus, g->table.data(), mode, &plan))
                  return group_fail(g);
            const spmv_b200_shard_desc &t = g->table[r];
            const int64_t M = t.r1 - t.r0, k0 = A->IRP[t.r0];
            std::vector<int> irp((size_t)M + 1);
            for (int64_t i = 0; i <= M; ++i)
                  irp[i] = A->IRP[t.r0 + i] - (int)k0;
            if (spmv_b200_set_device(r))
                  return group_fail(g);
            spmv_b200_csr *sh = spmv_b200_csr_create_ex(M, t.c1 - t.c0, irp[M], irp.data(), 4, A->JA + k0,
                                                        A->AS + k0, t.c0, plan.cuts, plan.n_cuts);
            if (!sh)
                  return group_fail(g);
            spmv_b200_dist *d = spmv_b200_dist_create(&plan, g->table.data(), sh, kernel, wpb);
            if (!d) {
                  spmv_b200_csr_destroy(sh);
                  return group_fail(g);
            }
            d->own_shard = true;
            g->ranks.push_back(d);
      }
      if (group_wire(g))
            return group_fail(g);
      spmv_b200_set_device(g->home_device);
      return g;
}

extern "C" spmv_b200_dist_group *spmv_b200_dist_group_stencil27(int nx, int ny, int nz, int n_gpus,
                                                                int kernel, int wpb, int mode) {
      if (ensure_device())
            return nullptr;
      if (n_gpus < 1 || n_gpus > SPMV_B200_MAX_RANKS || n_gpus > spmv_b200_device_count() ||
          nz < n_gpus) {
            fail(-EINVAL, "dist_group_stencil27: 1..%d GPUs (%d visible), at least one plane each",
                 SPMV_B200_MAX_RANKS, spmv_b200_device_count());
            return nullptr;
      }
      auto *g = new spmv_b200_dist_group();
      cudaGetDevice(&g->home_device);
      g->n = n_gpus, g->N = (int64_t)nx * ny * nz;
      g->table.resize(n_gpus);
      std::vector<int> z((size_t)n_gpus + 1);
      for (int r = 0; r <= n_gpus; ++r)
            z[r] = (int)((int64_t)nz * r / n_gpus);
      for (int r = 0; r < n_gpus; ++r)
            if (spmv_b200_stencil27_shard_desc(nx, ny, nz, z[r], z[r + 1], &g->table[r]))
                  return group_fail(g);
      for (int r = 0; r < n_gpus; ++r) {
            spmv_b200_dist_plan plan;
            if (spmv_b200_dist_make_plan(r, n_gpus, g->table.data(), mode, &plan))
                  return group_fail(g);
            const spmv_b200_shard_desc &t = g->table[r];
            if (spmv_b200_set_device(r))
                  return group_fail(g);
            spmv_b200_csr *sh = spmv_b200_csr_gen_stencil27(nx, ny, nz, z[r], z[r + 1], t.c0,
                                                            t.c1 - t.c0, plan.cuts, plan.n_cuts);
            if (!sh)
                  return group_fail(g);
            spmv_b200_dist *d = spmv_b200_dist_create(&plan, g->table.data(), sh, kernel, wpb);
            if (!d) {
                  spmv_b200_csr_destroy(sh);
                  return group_fail(g);
            }
            d->own_shard = true;
            g->ranks.push_back(d);
      }
      if (group_wire(g))
            return group_fail(g);
      spmv_b200_set_device(g->home_device);
      return g;
}

extern "C" int spmv_b200_dist_group_size(const spmv_b200_dist_group *g) { return g ? g->n : -1; }
extern "C" spmv_b200_dist *spmv_b200_dist_group_rank(spmv_b200_dist_group *g, int r) {
      return g && r >= 0 && r < g->n ? g->ranks[r] : nullptr;
}

extern "C" int spmv_b200_dist_group_set_x(spmv_b200_dist_group *g, const double *x_host) {
      if (!g || !x_host)
            return fail(-EINVAL, "dist_group_set_x: null argument");
      for (auto *d : g->ranks) { // the job must be idle before any halo is overwritten
            int rc = spmv_b200_dist_sync(d);
            if (rc)
                  return rc;
      }
      const bool nccl = g->ranks[0]->mode == SPMV_B200_DIST_NCCL && g->n > 1;
      if (nccl)
            B200_NCCL(g_nccl.GroupStart());
      for (auto *d : g->ranks) {
            int rc = spmv_b200_dist_set_x(d, x_host + d->me.r0);
            if (rc)
                  return rc;
      }
      if (nccl)
            B200_NCCL(g_nccl.GroupEnd());
      cudaSetDevice(g->home_device);
      return 0;
}

// k steps on every GPU; *ms_out (may be NULL) = the longest per-GPU time between CUDA events
// recorded around the k steps on that GPU's stream.  Returns after all streams have drained.
extern "C" int spmv_b200_dist_group_iterate(spmv_b200_dist_group *g, int k, double *ms_out) {
      if (!g || k < 0)
            return fail(-EINVAL, "dist_group_iterate: bad arguments");
      const bool nccl = g->ranks[0]->mode == SPMV_B200_DIST_NCCL && g->n > 1;
      for (int r = 0; r < g->n; ++r) {
            B200_CUDA(cudaSetDevice(g->ranks[r]->device));
            B200_CUDA(cudaEventRecord(g->ev0[r], g->ranks[r]->st));
      }
      if (!nccl) {
            // interleave the ranks in slices of two steps: every GPU gets work early, and no
            // launch queue fills up behind a wait kernel whose neighbour has not been fed yet
            for (int done = 0; done < k;) {
                  const int chunk = std::min(k - done, 2);
                  for (auto *d : g->ranks) {
                        int rc = dist_iterate(d, chunk);
                        if (rc)
                              return rc;
                  }
                  done += chunk;
            }
      } else {
            for (int s = 0; s < k; ++s) {
                  for (auto *d : g->ranks) {
                        B200_CUDA(cudaSetDevice(d->device));
                        int rc = dist_step(d, 1);
                        if (rc)
                              return rc;
                  }
                  B200_NCCL(g_nccl.GroupStart());
                  for (auto *d : g->ranks) {
                        B200_CUDA(cudaSetDevice(d->device));
                        int rc = dist_step(d, 2);
                        if (rc)
                              return rc;
                  }
                  B200_NCCL(g_nccl.GroupEnd());
                  for (auto *d : g->ranks) {
                        B200_CUDA(cudaSetDevice(d->device));
                        int rc = dist_step(d, 3);
                        if (rc)
                              return rc;
                  }
            }
      }
      for (int r = 0; r < g->n; ++r) {
            B200_CUDA(cudaSetDevice(g->ranks[r]->device));
            B200_CUDA(cudaEventRecord(g->ev1[r], g->ranks[r]->st));
      }
      double worst = 0.0;
      for (int r = 0; r < g->n; ++r) {
            int rc = spmv_b200_dist_sync(g->ranks[r]);
            if (rc)
                  return rc;
            float ms = 0.f;
            B200_CUDA(cudaEventElapsedTime(&ms, g->ev0[r], g->ev1[r]));
            worst = std::max(worst, (double)ms);
      }
      if (ms_out)
            *ms_out = worst;
      cudaSetDevice(g->home_device);
      return 0;
}

extern "C" int spmv_b200_dist_group_get_x(spmv_b200_dist_group *g, double *x_host) {
      if (!g || !x_host)
            return fail(-EINVAL, "dist_group_get_x: null argument");
      for (auto *d : g->ranks) {
            int rc = spmv_b200_dist_get_x(d, x_host + d->me.r0);
            if (rc)
                  return rc;
      }
      cudaSetDevice(g->home_device);
      return 0;
}

extern "C" void spmv_b200_dist_group_destroy(spmv_b200_dist_group *g) {
      if (!g)
            return;
      for (auto *d : g->ranks) // drain everything before any buffer a peer writes into goes away
            if (d) {
                  cudaSetDevice(d->device);
                  cudaStreamSynchronize(d->st);
                  cudaStreamSynchronize(d->comm);
            }
      for (size_t r = 0; r < g->ev0.size(); ++r) {
            cudaEventDestroy(g->ev0[r]);
            cudaEventDestroy(g->ev1[r]);
      }
      for (auto *d : g->ranks)
            spmv_b200_dist_destroy(d);
      cudaSetDevice(g->home_device);
      cudaGetLastError();
      delete g;
}
