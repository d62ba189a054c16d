// comm.cu -- multi-GPU behind the C ABI: communicators, the peer barrier, data-parallel training and tile-parallel
// rendering driven from C++ (SURVEY 8b `vnr_comm_init(n_devices)` + the same calls; 8e).
//
// What shards (and only that):
//   * rendering: rank r marches the image strips r, r + world, ... (march.cuh ray_to_pixel); finished pixels are stored by
//     the compositing kernels themselves -- into ONE pinned host frame shared by all ranks (every GPU delivers its strips over
//     its own PCIe link; vnr_map_frame on rank 0 returns that frame) or, with the download disabled, into rank 0's device
//     frame over NVLink.  A stream-ordered peer barrier closes the frame.  No collective, no copy through a staging buffer.
//   * training: synchronous data parallel.  Rank r draws the r-th of `world` consecutive batches of the ONE sampler stream,
//     runs forward + loss + backward with the loss normalised by the global batch, then ONE kernel per rank does
//     reduce-scatter + Adam + all-gather over peer memory (train.cu adam_grid_sharded_kernel) between two peer barriers.
//     Macrocell value ranges are merged (min / max) once per vnr_volume_train call -- ranges only ever grow, so merging at
//     the end equals merging every step.
// Everything else is replicated.
#include <atomic>
#include <chrono>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>

#include "comm.h"
#include "render.h"
#include "train.h"

namespace vnr {

// ------------------------------------------------------------------------------------------------------------------
// peer barrier
// ------------------------------------------------------------------------------------------------------------------
struct PeerBarrierArgs { unsigned long long* peer[kMaxPeers]; };

__global__ void peer_barrier_kernel(PeerBarrierArgs a, unsigned long long* local, unsigned long long* h_flag, int rank, unsigned long long epoch) {
  const int t = threadIdx.x;
  __threadfence_system();                            // everything this stream did before is visible to the peers
  if (t != rank) {
    *reinterpret_cast<volatile unsigned long long*>(a.peer[t] + rank) = epoch;
    unsigned long long t0 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*reinterpret_cast<volatile unsigned long long*>(local + t) < epoch) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (now - t0 > 5000000000ull) {              // a peer never arrived: flag it (device word + pinned host word the host checks) and go on
        *reinterpret_cast<volatile unsigned long long*>(local + kMaxPeers) = epoch;
        *reinterpret_cast<volatile unsigned long long*>(h_flag) = epoch;
        break;
      }
    }
  }
  __threadfence_system();                            // what the peers published before their flag is visible after this kernel
}

PeerBarrier::~PeerBarrier() {
  if (ipc) for (int r = 0; r < world; ++r) if (r != rank && peer[r]) cudaIpcCloseMemHandle(peer[r]);
  if (local) cudaFree(local);
  if (h_flag) cudaFreeHost(h_flag);
}

PeerBarrier* peer_barrier_create() {
  std::unique_ptr<PeerBarrier> b(new PeerBarrier());
  VNR_CUDA(cudaMalloc((void**)&b->local, sizeof(unsigned long long) * (kMaxPeers + 1)));
  VNR_CUDA(cudaMemset(b->local, 0, sizeof(unsigned long long) * (kMaxPeers + 1)));
  VNR_CUDA(cudaMallocHost((void**)&b->h_flag, sizeof(unsigned long long)));
  *b->h_flag = 0;
  return b.release();
}

void peer_barrier_attach_ipc(PeerBarrier* b, int rank, int world, const void* all_handles) {
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || (world > 1 && !all_handles)) throw InvalidError("bad barrier rank / world");
  b->rank = rank; b->world = world; b->ipc = true;
  for (int r = 0; r < world; ++r) {
    if (r == rank) { b->peer[r] = b->local; continue; }
    cudaIpcMemHandle_t h; memcpy(&h, reinterpret_cast<const char*>(all_handles) + (size_t)r * 64, 64);
    VNR_CUDA(cudaIpcOpenMemHandle((void**)&b->peer[r], h, cudaIpcMemLazyEnablePeerAccess));
  }
}

void peer_barrier_attach_ptrs(PeerBarrier* b, int rank, int world, unsigned long long* const* flags) {
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) throw InvalidError("bad barrier rank / world");
  b->rank = rank; b->world = world; b->ipc = false;
  for (int r = 0; r < world; ++r) b->peer[r] = r == rank ? b->local : flags[r];
}

void peer_barrier_sync(PeerBarrier* b, cudaStream_t s) {
  if (b->world <= 1) return;
  PeerBarrierArgs a;
  for (int r = 0; r < kMaxPeers; ++r) a.peer[r] = b->peer[r];
  ++b->epoch;
  peer_barrier_kernel<<<1, b->world, 0, s>>>(a, b->local, b->h_flag, b->rank, b->epoch);
  VNR_CUDA(cudaGetLastError());
}

// epoch of the last barrier that gave up waiting (0 = healthy).  Reads the pinned host word: valid for every barrier kernel
// that has completed, so callers check after a synchronisation they do anyway (mapping a frame, reading the loss).
unsigned long long peer_barrier_timed_out(PeerBarrier* b) { return *reinterpret_cast<volatile unsigned long long*>(b->h_flag); }

void peer_barrier_require_healthy(PeerBarrier* b, const char* what) {
  if (!b || b->world <= 1) return;
  const unsigned long long e = peer_barrier_timed_out(b);
  if (e) throw StateError(std::string(what) + ": a rank did not reach the peer barrier within 5 s (barrier call " + std::to_string(e) +
                          "); its contribution is missing -- the result is not valid");
}

// ------------------------------------------------------------------------------------------------------------------
// communicators
// ------------------------------------------------------------------------------------------------------------------
constexpr uint32_t kBoardMagic = 0x564e5243u;        // "VNRC"

struct CommBoard {
  std::atomic<uint32_t> magic;
  uint32_t world;
  std::atomic<uint32_t> arrived;                     // total arrivals at the host barrier so far
  uint32_t pad_;
  uint8_t payload[kMaxPeers][kCommPayload];
};

struct CommLocal {
  int world = 1;
  std::vector<Comm*> comms;
  std::map<uint32_t, std::array<Volume*, kMaxPeers>> vols;
  std::map<uint32_t, std::array<Renderer*, kMaxPeers>> rens;
};

static void* shm_map(const std::string& name, size_t bytes, bool create, int timeout_s = 120) {
  const std::string path = "/" + name;
  int fd = -1;
  if (create) {
    shm_unlink(path.c_str());
    fd = shm_open(path.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) throw InvalidError("cannot create the shared-memory segment " + path);
    if (ftruncate(fd, (off_t)bytes) != 0) { close(fd); throw InvalidError("cannot size the shared-memory segment " + path); }
  } else {
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
      fd = shm_open(path.c_str(), O_RDWR, 0600);
      if (fd >= 0) {
        struct stat st;
        if (fstat(fd, &st) == 0 && (size_t)st.st_size >= bytes) break;
        close(fd); fd = -1;
      }
      if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(timeout_s)) throw StateError("timed out waiting for the shared-memory segment " + path + " (is rank 0 running?)");
      std::this_thread::sleep_for(std::chrono::milliseconds(2));
    }
  }
  void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) throw InvalidError("mmap of " + path + " failed");
  return p;
}

Comm::~Comm() {
  if (board) {
    munmap(board, sizeof(CommBoard));
    if (rank == 0) shm_unlink(("/" + name).c_str());
  }
}

void Comm::host_barrier() {
  if (world <= 1 || in_process) return;
  const uint64_t target = (++bar_calls) * (uint64_t)world;
  board->arrived.fetch_add(1, std::memory_order_acq_rel);
  const auto t0 = std::chrono::steady_clock::now();
  uint32_t spins = 0;
  while ((uint64_t)board->arrived.load(std::memory_order_acquire) < target) {
    if (++spins > 2000) std::this_thread::sleep_for(std::chrono::microseconds(50));
    if ((spins & 0xFFFu) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) throw StateError("communicator barrier timed out: a rank is missing");
  }
}

void Comm::allgather(const void* mine, size_t bytes, void* all) {
  if (bytes > kCommPayload) throw InvalidError("communicator payload too large");
  if (world <= 1 || in_process) { memcpy(all, mine, bytes); return; }
  memcpy(board->payload[rank], mine, bytes);
  host_barrier();
  for (int r = 0; r < world; ++r) memcpy((char*)all + (size_t)r * bytes, board->payload[r], bytes);
  host_barrier();                                    // nobody overwrites its payload before everybody has read it
}

Comm* comm_create_rank(int rank, int world, const char* name) {
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) throw InvalidError("bad communicator rank / world (at most 8 ranks of one NVSwitch box)");
  if (world > 1 && (!name || !*name)) throw InvalidError("a rendezvous name is required");
  std::unique_ptr<Comm> c(new Comm());
  c->rank = rank; c->world = world; c->name = name ? name : "";
  VNR_CUDA(cudaGetDevice(&c->device));
  if (world > 1) {
    c->board = reinterpret_cast<CommBoard*>(shm_map(c->name, sizeof(CommBoard), rank == 0));
    if (rank == 0) { c->board->world = (uint32_t)world; c->board->arrived.store(0); c->board->magic.store(kBoardMagic, std::memory_order_release); }
    else {
      const auto t0 = std::chrono::steady_clock::now();
      while (c->board->magic.load(std::memory_order_acquire) != kBoardMagic) {
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(120)) throw StateError("timed out waiting for rank 0 to publish the communicator board");
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
      }
      if (c->board->world != (uint32_t)world) throw InvalidError("communicator world size differs between ranks");
    }
    c->host_barrier();
  }
  return c.release();
}

std::vector<Comm*> comm_create_local(int n) {
  int have = 0;
  VNR_CUDA(cudaGetDeviceCount(&have));
  // testing knob: VNR_COMM_SHARE_DEVICES=1 lets several ranks share a device (rank r -> device r % visible), so the whole
  // multi-rank control and data path runs on a one-GPU box (the barrier kernels of the ranks are co-resident)
  const char* share_env = getenv("VNR_COMM_SHARE_DEVICES");
  const bool share = share_env && atoi(share_env) != 0;
  if (n < 1 || n > kMaxPeers || have < 1 || (n > have && !share))
    throw InvalidError("vnr_comm_init: " + std::to_string(n) + " devices requested, " + std::to_string(have) + " visible (at most 8)");
  int before = 0;
  VNR_CUDA(cudaGetDevice(&before));
  const int n_dev = n < have ? n : have;
  for (int i = 0; i < n_dev; ++i) {
    VNR_CUDA(cudaSetDevice(i));
    for (int j = 0; j < n_dev; ++j) {
      if (i == j) continue;
      int can = 0;
      VNR_CUDA(cudaDeviceCanAccessPeer(&can, i, j));
      if (!can) throw UnsupportedError("devices " + std::to_string(i) + " and " + std::to_string(j) + " have no peer access");
      const cudaError_t e = cudaDeviceEnablePeerAccess(j, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) VNR_CUDA(e);
      cudaGetLastError();
    }
  }
  VNR_CUDA(cudaSetDevice(before));
  auto local = std::make_shared<CommLocal>();
  local->world = n;
  std::vector<Comm*> out;
  for (int r = 0; r < n; ++r) {
    Comm* c = new Comm();
    c->rank = r; c->world = n; c->device = r % n_dev; c->in_process = true; c->local = local;
    out.push_back(c);
  }
  local->comms = out;
  return out;
}

// ------------------------------------------------------------------------------------------------------------------
// data-parallel training
// ------------------------------------------------------------------------------------------------------------------
VolumeComm::~VolumeComm() {
  for (void* p : ipc_open) cudaIpcCloseMemHandle(p);
  delete barrier;
}

__global__ void master_from_params_kernel(size_t n, const __half* __restrict__ p, float* __restrict__ m) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) m[i] = __half2float(p[i]);
}

struct McMergeArgs { const float* range[kMaxPeers]; };
__global__ void mc_merge_kernel(McMergeArgs a, int world, size_t n2, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2) return;
  float v = a.range[0][i];
  for (int r = 1; r < world; ++r) { const float w = a.range[r][i]; v = (i & 1) ? fmaxf(v, w) : fminf(v, w); }     // (min - 1, max + 1) pairs
  out[i] = v;
}

struct VolumeExport {           // what one rank contributes to the exchange (IPC handles or plain pointers)
  cudaIpcMemHandle_t params, grid_grads, mlp_grads, mc_range, loss, flags;
  void* ptr[6];
  Pcg32 rng;
};

// After the peers are known: every replica starts from rank 0's parameters, macrocell ranges and sampler stream, with a
// fresh optimizer (fp32 master = the fp16 blob, as Trainer::set_params does, trainer.h:281-297).
static void sync_replica_begin(Volume* v, const Pcg32& rng0) {
  VolumeComm* vc = v->vcomm;
  cudaStream_t s = v->stream;
  if (vc->comm->rank != 0) {
    VNR_CUDA(cudaMemcpyAsync(v->params.p, v->dp_params[0], v->params.bytes(), cudaMemcpyDefault, s));
    VNR_CUDA(cudaMemcpyAsync(v->mc_range.p, vc->mc_range[0], v->mc_range.bytes(), cudaMemcpyDefault, s));
  }
  v->sampler_rng = rng0;
  const size_t n = v->cfg.n_params();
  v->master.alloc(n);
  master_from_params_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, v->params.p, v->master.p);
  VNR_CUDA(cudaGetLastError());
  v->have_params = true;
  reset_optimizer_state(v);                          // synchronises the stream
  if (v->n_alpha > 0) { macrocell_update_max_opacity(v, s); VNR_CUDA(cudaStreamSynchronize(s)); }
}

static void fill_export(Volume* v, VolumeExport& e, bool ipc) {
  VolumeComm* vc = v->vcomm;
  memset(&e, 0, sizeof e);
  e.ptr[0] = v->params.p; e.ptr[1] = v->grid_grads.p; e.ptr[2] = v->mlp_grads.p; e.ptr[3] = v->mc_range.p; e.ptr[4] = v->loss_accum.p; e.ptr[5] = vc->barrier->local;
  e.rng = v->sampler_rng;
  if (ipc) {
    VNR_CUDA(cudaIpcGetMemHandle(&e.params, v->params.p));
    VNR_CUDA(cudaIpcGetMemHandle(&e.grid_grads, v->grid_grads.p));
    VNR_CUDA(cudaIpcGetMemHandle(&e.mlp_grads, v->mlp_grads.p));
    VNR_CUDA(cudaIpcGetMemHandle(&e.mc_range, v->mc_range.p));
    VNR_CUDA(cudaIpcGetMemHandle(&e.loss, v->loss_accum.p));
    VNR_CUDA(cudaIpcGetMemHandle(&e.flags, vc->barrier->local));
  }
}

static void wire_volume(Volume* v, const VolumeExport* all, bool ipc) {
  VolumeComm* vc = v->vcomm;
  const int R = vc->comm->rank, W = vc->comm->world;
  v->dp_rank = R; v->dp_world = W;
  unsigned long long* flags[kMaxPeers] = {};
  for (int r = 0; r < W; ++r) {
    if (r == R) {
      v->dp_params[r] = v->params.p; v->dp_grid_grads[r] = v->grid_grads.p; v->dp_mlp_grads[r] = v->mlp_grads.p;
      vc->mc_range[r] = v->mc_range.p; vc->loss_accum[r] = v->loss_accum.p; flags[r] = vc->barrier->local;
      continue;
    }
    if (ipc) {
      void* p[6];
      const cudaIpcMemHandle_t* h[6] = {&all[r].params, &all[r].grid_grads, &all[r].mlp_grads, &all[r].mc_range, &all[r].loss, &all[r].flags};
      for (int k = 0; k < 6; ++k) { VNR_CUDA(cudaIpcOpenMemHandle(&p[k], *h[k], cudaIpcMemLazyEnablePeerAccess)); if (k < 5) vc->ipc_open.push_back(p[k]); }
      v->dp_params[r] = p[0]; v->dp_grid_grads[r] = p[1]; v->dp_mlp_grads[r] = p[2];
      vc->mc_range[r] = (float*)p[3]; vc->loss_accum[r] = (double*)p[4]; flags[r] = (unsigned long long*)p[5];
    } else {
      v->dp_params[r] = all[r].ptr[0]; v->dp_grid_grads[r] = all[r].ptr[1]; v->dp_mlp_grads[r] = all[r].ptr[2];
      vc->mc_range[r] = (float*)all[r].ptr[3]; vc->loss_accum[r] = (double*)all[r].ptr[4]; flags[r] = (unsigned long long*)all[r].ptr[5];
    }
  }
  if (ipc) {      // the barrier owns (and closes) its mappings
    vc->barrier->rank = R; vc->barrier->world = W; vc->barrier->ipc = true;
    for (int r = 0; r < W; ++r) vc->barrier->peer[r] = flags[r];
  } else peer_barrier_attach_ptrs(vc->barrier, R, W, flags);
  vc->mc_merged.alloc(v->mc_range.n);
  vc->resolved = true;
  outofcore_set_rank(v, R);                          // out-of-core ground truth: every rank keeps its own random slabs
}

void comm_attach_volume(Volume* v, Comm* c) {
  if (!c) throw InvalidError("null communicator");
  if (v->vcomm) throw StateError("the volume is already attached to a communicator");
  if (v->dp_world) throw StateError("detach the data-parallel peers (vnr_volume_dp_detach) first");
  VNR_CUDA(cudaStreamSynchronize(v->stream));
  const bool had_params = v->have_params;
  if (!had_params && c->rank == 0) throw StateError("rank 0 attaches a volume whose parameters are set (they are replicated to the other ranks)");
  v->have_params = true;                             // ranks != 0 receive rank 0's parameters below
  if (!v->have_opt) reset_optimizer_state(v);
  train_ensure_buffers(v);
  {   // every kernel of a data-parallel step is loaded before the first peer barrier exists (see train_preload_kernels)
    cudaFuncAttributes fa;
    VNR_CUDA(cudaFuncGetAttributes(&fa, peer_barrier_kernel));
    VNR_CUDA(cudaFuncGetAttributes(&fa, mc_merge_kernel));
    VNR_CUDA(cudaFuncGetAttributes(&fa, master_from_params_kernel));
    train_preload_kernels(v);
  }
  std::unique_ptr<VolumeComm> vc(new VolumeComm());
  vc->comm = c; vc->id = c->n_volumes++;
  vc->barrier = peer_barrier_create();
  v->vcomm = vc.release();
  VolumeExport mine;
  if (c->world == 1) {
    fill_export(v, mine, false);
    wire_volume(v, &mine, false);
    return;
  }
  if (c->in_process) {
    auto& slot = c->local->vols[v->vcomm->id];
    slot[(size_t)c->rank] = v;
    for (int r = 0; r < c->world; ++r) if (!slot[(size_t)r]) return;           // resolved when the last rank attaches
    VolumeExport all[kMaxPeers];
    int before = 0; VNR_CUDA(cudaGetDevice(&before));
    for (int r = 0; r < c->world; ++r) { VNR_CUDA(cudaSetDevice(slot[(size_t)r]->device)); fill_export(slot[(size_t)r], all[r], false); }
    for (int r = 0; r < c->world; ++r) { VNR_CUDA(cudaSetDevice(slot[(size_t)r]->device)); wire_volume(slot[(size_t)r], all, false); }
    for (int r = 0; r < c->world; ++r) { VNR_CUDA(cudaSetDevice(slot[(size_t)r]->device)); sync_replica_begin(slot[(size_t)r], all[0].rng); }
    VNR_CUDA(cudaSetDevice(before));
    return;
  }
  fill_export(v, mine, true);
  VolumeExport all[kMaxPeers];
  c->allgather(&mine, sizeof mine, all);
  wire_volume(v, all, true);
  sync_replica_begin(v, all[0].rng);
  c->host_barrier();                                 // every rank has copied rank 0's blob before anybody trains
}

void comm_detach_volume(Volume* v) {
  if (!v->vcomm) return;
  VNR_CUDA(cudaStreamSynchronize(v->stream));
  Comm* c = v->vcomm->comm;
  if (c->in_process && c->local) {
    auto it = c->local->vols.find(v->vcomm->id);
    if (it != c->local->vols.end()) it->second[(size_t)c->rank] = nullptr;
  }
  if (!c->in_process && c->world > 1) {
    for (int r = 0; r < v->dp_world; ++r) {
      if (r == v->dp_rank) continue;
      // params / gradient mappings are in ipc_open; the barrier closes its own
    }
  }
  for (int r = 0; r < kMaxPeers; ++r) v->dp_params[r] = v->dp_grid_grads[r] = v->dp_mlp_grads[r] = nullptr;
  v->dp_world = 0; v->dp_rank = 0;
  delete v->vcomm; v->vcomm = nullptr;
}

// vnrNeuralVolumeTrain across the ranks of the communicator (NeuralVolume::Impl::train, network.cu:231-259, per rank)
void comm_train_steps(Volume* v, int steps, size_t batch, bool update_macrocell, cudaStream_t s) {
  VolumeComm* vc = v->vcomm;
  if (!vc || !vc->resolved) throw StateError("every rank must attach its volume to the communicator before training");
  const int R = vc->comm->rank, W = vc->comm->world;
  if (!v->have_gt && !v->ooc) throw StateError("[error]: missing a reference volume.");
  if (batch == 0) batch = 1 << 16;                                       // network.cu:183
  if (batch % kTile) throw InvalidError("Batch size must be a multiple of 128.");
  v->train_x.ensure(3 * batch); v->train_y.ensure(batch);
  if (steps <= 0) return;                                                // nothing is drawn, nothing to merge
  const uint64_t ups = v->ooc ? 5 : 3;                                   // uniforms per sample of the sampler in use
  // rank r takes the r-th of `world` consecutive batches of the one sampler stream
  auto draw = [&](float* x, float* y, cudaStream_t st) {
    v->sampler_rng.advance((uint64_t)R * ups * batch);
    sample_batch(v, x, y, batch, st);
    v->sampler_rng.advance((uint64_t)(W - 1 - R) * ups * batch);
  };
  if (v->ooc || getenv("VNR_TRAIN_SERIAL")) {                            // out-of-core batches come through pinned staging buffers in stream order
    for (int i = 0; i < steps; ++i) {
      draw(v->train_x.p, v->train_y.p, s);
      train_grads(v, v->train_x.p, v->train_y.p, batch, batch * (size_t)W, s);
      peer_barrier_sync(vc->barrier, s);                                 // every rank's gradients are complete
      dp_optimizer_step(v, s);                                           // reduce-scatter + Adam + all-gather over peer memory
      peer_barrier_sync(vc->barrier, s);                                 // every rank's parameters are complete
      dp_finish_step(v, s);
      if (update_macrocell) macrocell_update_explicit(v, v->train_x.p, v->train_y.p, batch, s);
    }
  } else {
    // as train_steps (train.cu): the macrocell update of this batch and the draw of the NEXT batch run on the volume's
    // high-priority side stream under the barriers and the optimizer; nothing of it touches peer memory
    train_side_stream(v);
    v->train_x2.ensure(3 * batch); v->train_y2.ensure(batch);
    float* xb[2] = {v->train_x.p, v->train_x2.p};
    float* yb[2] = {v->train_y.p, v->train_y2.p};
    draw(xb[0], yb[0], s);
    for (int i = 0; i < steps; ++i) {
      const int b = i & 1;
      train_grads(v, xb[b], yb[b], batch, batch * (size_t)W, s);
      VNR_CUDA(cudaEventRecord(v->ev_fork, s));
      VNR_CUDA(cudaStreamWaitEvent(v->side, v->ev_fork, 0));
      if (update_macrocell) macrocell_update_explicit(v, xb[b], yb[b], batch, v->side);
      if (i + 1 < steps) draw(xb[b ^ 1], yb[b ^ 1], v->side);
      VNR_CUDA(cudaEventRecord(v->ev_join, v->side));
      peer_barrier_sync(vc->barrier, s);                                 // every rank's gradients are complete
      dp_optimizer_step(v, s);                                           // reduce-scatter + Adam + all-gather over peer memory
      peer_barrier_sync(vc->barrier, s);                                 // every rank's parameters are complete
      dp_finish_step(v, s);
      VNR_CUDA(cudaStreamWaitEvent(s, v->ev_join, 0));
    }
  }
  if (update_macrocell && steps > 0 && W > 1) {
    McMergeArgs a;
    for (int r = 0; r < kMaxPeers; ++r) a.range[r] = r < W ? vc->mc_range[r] : nullptr;
    const size_t n2 = v->mc_range.n;
    peer_barrier_sync(vc->barrier, s);                                   // every rank's ranges include its last batch
    mc_merge_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, s>>>(a, W, n2, vc->mc_merged.p);
    VNR_CUDA(cudaGetLastError());
    peer_barrier_sync(vc->barrier, s);                                   // everybody has read everybody's ranges
    VNR_CUDA(cudaMemcpyAsync(v->mc_range.p, vc->mc_merged.p, v->mc_range.bytes(), cudaMemcpyDeviceToDevice, s));
  }
}

// sum over ranks of loss accumulator `which` (0: running sum over steps, 1: last step); the local stream is synchronised,
// and the two barriers of a step order every peer's accumulation before it
double comm_global_loss(Volume* v, int which) {
  VolumeComm* vc = v->vcomm;
  VNR_CUDA(cudaStreamSynchronize(v->stream));
  peer_barrier_require_healthy(vc->barrier, "data-parallel training");
  double total = 0;
  for (int r = 0; r < vc->comm->world; ++r) {
    double x = 0;
    VNR_CUDA(cudaMemcpy(&x, vc->loss_accum[r] + which, sizeof x, cudaMemcpyDefault));
    total += x;
  }
  return total;
}

// ------------------------------------------------------------------------------------------------------------------
// tile-parallel rendering
// ------------------------------------------------------------------------------------------------------------------
RendererComm::~RendererComm() {
  for (PeerBarrier* b : barriers) delete b;
  for (void* p : ipc_open) cudaIpcCloseMemHandle(p);
  if (host_base) {
    if (host_is_shm) {
      cudaHostUnregister(host_base);
      munmap(host_base, host_bytes);
      if (comm && comm->rank == 0) shm_unlink(("/" + shm_name).c_str());
    } else if (comm && comm->rank == 0) cudaFreeHost(host_base);
  }
}

struct RendererExport {
  cudaIpcMemHandle_t frame[kMaxFramesInFlight], flags[kMaxFramesInFlight];
  void* frame_ptr[kMaxFramesInFlight]; void* flags_ptr[kMaxFramesInFlight];
  void* host_base;
  int width, height, n_slots;
};

static void fill_export(Renderer* r, RendererExport& e, bool ipc) {
  RendererComm* rc = r->rcomm;
  memset(&e, 0, sizeof e);
  e.width = r->width; e.height = r->height; e.n_slots = (int)r->slots.size();
  e.host_base = rc->host_base;
  for (int k = 0; k < e.n_slots; ++k) {
    e.frame_ptr[k] = r->slot(k).frame.p; e.flags_ptr[k] = rc->barriers[(size_t)k]->local;
    if (ipc) {
      VNR_CUDA(cudaIpcGetMemHandle(&e.frame[k], r->slot(k).frame.p));
      VNR_CUDA(cudaIpcGetMemHandle(&e.flags[k], rc->barriers[(size_t)k]->local));
    }
  }
}

static void wire_renderer(Renderer* r, const RendererExport* all, bool ipc) {
  RendererComm* rc = r->rcomm;
  const int R = rc->comm->rank, W = rc->comm->world;
  for (int q = 0; q < W; ++q)
    if (all[q].width != r->width || all[q].height != r->height || all[q].n_slots != (int)r->slots.size())
      throw InvalidError("the ranks attach renderers of different frame size / frames in flight");
  const size_t npix = (size_t)r->width * r->height;
  for (int k = 0; k < (int)r->slots.size(); ++k) {
    FrameSlot& S = r->slot(k);
    PeerBarrier* b = rc->barriers[(size_t)k];
    unsigned long long* flags[kMaxPeers] = {};
    for (int q = 0; q < W; ++q) {
      if (q == R) { flags[q] = b->local; continue; }
      if (ipc) { void* p; VNR_CUDA(cudaIpcOpenMemHandle(&p, all[q].flags[k], cudaIpcMemLazyEnablePeerAccess)); flags[q] = (unsigned long long*)p; }
      else flags[q] = (unsigned long long*)all[q].flags_ptr[k];
    }
    if (ipc) { b->rank = R; b->world = W; b->ipc = true; for (int q = 0; q < W; ++q) b->peer[q] = flags[q]; }
    else peer_barrier_attach_ptrs(b, R, W, flags);
    // rank 0's device frame of the slot: where the pixels go when frames stay on the device
    if (R != 0) {
      void* p = all[0].frame_ptr[k];
      if (ipc) { VNR_CUDA(cudaIpcOpenMemHandle(&p, all[0].frame[k], cudaIpcMemLazyEnablePeerAccess)); rc->ipc_open.push_back(p); }
      S.frame_target = reinterpret_cast<float4*>(p);
    }
    // the shared pinned host frames of the slot
    for (int h = 0; h < 2; ++h) {
      if (S.h_frame[h] && !S.h_frame_external) cudaFreeHost(S.h_frame[h]);
      S.h_frame[h] = reinterpret_cast<float4*>(rc->host_base) + ((size_t)k * 2 + (size_t)h) * npix;
    }
    S.h_frame_external = true;
    S.host_nonzero_valid[0] = S.host_nonzero_valid[1] = false;      // other frames: what they hold is unknown to this slot's masks
    S.rendered = false; S.downloaded = false; S.mapped = true;
  }
  r->part_rank = R; r->part_world = W;
  r->n_rendered = r->n_mapped = 0; r->last_slot = 0;
  r->reset = true;
  rc->width = r->width; rc->height = r->height; rc->n_slots = (int)r->slots.size();
  rc->resolved = true;
}

void comm_attach_renderer(Renderer* r, Comm* c) {
  if (!c) throw InvalidError("null communicator");
  if (r->rcomm) throw StateError("the renderer is already attached to a communicator");
  if (r->width <= 0 || r->height <= 0) throw StateError("set the framebuffer size (and the frames in flight) before attaching the renderer");
  r->sync_all();
  std::unique_ptr<RendererComm> rc(new RendererComm());
  rc->comm = c; rc->id = c->n_renderers++;
  for (size_t k = 0; k < r->slots.size(); ++k) rc->barriers.push_back(peer_barrier_create());
  const size_t npix = (size_t)r->width * r->height;
  rc->host_bytes = r->slots.size() * 2 * npix * sizeof(float4);
  r->rcomm = rc.release();
  RendererComm* q = r->rcomm;
  if (c->world == 1) {
    VNR_CUDA(cudaHostAlloc(&q->host_base, q->host_bytes, cudaHostAllocPortable | cudaHostAllocMapped));
    RendererExport mine; fill_export(r, mine, false);
    wire_renderer(r, &mine, false);
    return;
  }
  if (c->in_process) {
    auto& slot = c->local->rens[q->id];
    slot[(size_t)c->rank] = r;
    for (int k = 0; k < c->world; ++k) if (!slot[(size_t)k]) return;            // resolved when the last rank attaches
    int before = 0; VNR_CUDA(cudaGetDevice(&before));
    void* base = nullptr;
    VNR_CUDA(cudaHostAlloc(&base, q->host_bytes, cudaHostAllocPortable | cudaHostAllocMapped));
    RendererExport all[kMaxPeers];
    for (int k = 0; k < c->world; ++k) { slot[(size_t)k]->rcomm->host_base = base; VNR_CUDA(cudaSetDevice(slot[(size_t)k]->vol->device)); fill_export(slot[(size_t)k], all[k], false); }
    for (int k = 0; k < c->world; ++k) { VNR_CUDA(cudaSetDevice(slot[(size_t)k]->vol->device)); wire_renderer(slot[(size_t)k], all, false); }
    VNR_CUDA(cudaSetDevice(before));
    return;
  }
  // one process per GPU: the host frames live in a shared-memory segment that every rank registers with its CUDA context
  q->host_is_shm = true;
  q->shm_name = c->name + "-fb" + std::to_string(q->id);
  if (c->rank == 0) q->host_base = shm_map(q->shm_name, q->host_bytes, true);
  c->host_barrier();
  if (c->rank != 0) q->host_base = shm_map(q->shm_name, q->host_bytes, false);
  VNR_CUDA(cudaHostRegister(q->host_base, q->host_bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
  RendererExport mine; fill_export(r, mine, true);
  RendererExport all[kMaxPeers];
  c->allgather(&mine, sizeof mine, all);
  wire_renderer(r, all, true);
  c->host_barrier();
}

void comm_detach_renderer(Renderer* r) {
  if (!r->rcomm) return;
  r->sync_all();
  Comm* c = r->rcomm->comm;
  if (c->in_process && c->local) {
    auto it = c->local->rens.find(r->rcomm->id);
    if (it != c->local->rens.end()) it->second[(size_t)c->rank] = nullptr;
  }
  const size_t npix = (size_t)r->width * r->height;
  for (auto& sp : r->slots) {
    FrameSlot& S = *sp;
    S.frame_target = nullptr;
    if (S.h_frame_external) {
      for (int h = 0; h < 2; ++h) {
        S.h_frame[h] = nullptr;
        VNR_CUDA(cudaMallocHost((void**)&S.h_frame[h], npix * sizeof(float4)));
        memset(S.h_frame[h], 0, npix * sizeof(float4));
      }
      S.h_frame_external = false;
      S.host_nonzero_valid[0] = S.host_nonzero_valid[1] = false;
    }
    S.rendered = false; S.downloaded = false; S.mapped = true;
  }
  r->part_rank = 0; r->part_world = 1; r->n_rendered = r->n_mapped = 0; r->reset = true;
  delete r->rcomm; r->rcomm = nullptr;
}

}  // namespace vnr
