// comm.h -- multi-GPU plumbing behind the C ABI (one NVSwitch box): communicator, peer barrier, and the state a
// communicator-attached volume / renderer carries.
//
// The reference runs on one GPU; SURVEY 8(b) asks for `vnr_comm_init(n_devices)` + the same calls.  Two ways to span GPUs:
//   * one process per GPU (torchrun / mpirun): vnr_comm_init_rank(rank, world, name) -- the ranks meet in a POSIX
//     shared-memory segment `name` (a host barrier and an all-gather of small payloads); device buffers are shared through
//     CUDA IPC handles;
//   * one process, n devices: vnr_comm_init(n, comms[]) -- peer access between the devices, plain pointers.
// Either way the DATA plane never goes through the host or a collective library: gradients / parameters / pixels move by
// peer loads and stores issued from the product's own kernels over NVLink, ordered by a stream-ordered barrier kernel.
#pragma once
#include <array>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "volume.h"

namespace vnr {

struct Renderer;

// Stream-ordered barrier between the ranks over peer memory: one `world`-thread kernel publishes this rank's epoch into
// every peer's flag array and waits for theirs (system-scope release / acquire).
struct PeerBarrier {
  int rank = 0, world = 1;
  bool ipc = true;                                   // peers' flag arrays are IPC mappings (closed on release)
  unsigned long long epoch = 0;
  unsigned long long* local = nullptr;               // [kMaxPeers + 1]: slot r = rank r's last epoch; slot kMaxPeers = timeout flag
  unsigned long long* peer[kMaxPeers] = {};
  unsigned long long* h_flag = nullptr;              // pinned host word: epoch of the last barrier that timed out (0 = healthy)
  ~PeerBarrier();
};
PeerBarrier* peer_barrier_create();
void peer_barrier_attach_ipc(PeerBarrier* b, int rank, int world, const void* all_handles64);
void peer_barrier_attach_ptrs(PeerBarrier* b, int rank, int world, unsigned long long* const* flags);
void peer_barrier_sync(PeerBarrier* b, cudaStream_t s);
unsigned long long peer_barrier_timed_out(PeerBarrier* b);
void peer_barrier_require_healthy(PeerBarrier* b, const char* what);   // throws StateError after a timed-out barrier

struct CommBoard;      // shared-memory rendezvous board (one process per GPU)
struct CommLocal;      // in-process registry (one process, n devices)

struct Comm {
  int rank = 0, world = 1, device = 0;
  bool in_process = false;
  std::string name;
  CommBoard* board = nullptr;
  std::shared_ptr<CommLocal> local;
  uint32_t n_volumes = 0, n_renderers = 0;           // collective sequence numbers: the k-th attach on every rank forms group k
  uint64_t bar_calls = 0;
  ~Comm();
  void host_barrier();                               // one process per GPU only
  void allgather(const void* mine, size_t bytes, void* all);
};

constexpr size_t kCommPayload = 4096;

// what a communicator-attached volume adds (data-parallel training: comm.cu comm_train_steps)
struct VolumeComm {
  Comm* comm = nullptr;
  uint32_t id = 0;
  bool resolved = false;
  PeerBarrier* barrier = nullptr;
  float* mc_range[kMaxPeers] = {};                   // every rank's macrocell value ranges (merged after a train call)
  double* loss_accum[kMaxPeers] = {};                // every rank's loss accumulators (global loss = sum over ranks)
  DevBuf<float> mc_merged;
  std::vector<void*> ipc_open;                       // mappings to close on detach
  ~VolumeComm();
};

// what a communicator-attached renderer adds (tile-parallel rendering; frames are gathered on rank 0)
struct RendererComm {
  Comm* comm = nullptr;
  uint32_t id = 0;
  bool resolved = false;
  int width = 0, height = 0, n_slots = 0;            // geometry the exchange was made for
  std::vector<PeerBarrier*> barriers;                // one per frame slot
  void* host_base = nullptr; size_t host_bytes = 0;  // shared pinned host frames: [slot][2][w*h] float4
  bool host_is_shm = false; std::string shm_name;
  std::vector<void*> ipc_open;
  ~RendererComm();
};

Comm* comm_create_rank(int rank, int world, const char* name);
std::vector<Comm*> comm_create_local(int n_devices);
void comm_attach_volume(Volume* v, Comm* c);
void comm_detach_volume(Volume* v);
void comm_train_steps(Volume* v, int steps, size_t batch, bool update_macrocell, cudaStream_t s);
double comm_global_loss(Volume* v, int which);       // which 0: running sum over steps, 1: last step
void comm_attach_renderer(Renderer* r, Comm* c);
void comm_detach_renderer(Renderer* r);

}  // namespace vnr
