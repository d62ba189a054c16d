// vnr_c_render.inl -- renderer entry points (MainRenderer / api.cpp:419-525)
static Renderer* R(vnr_renderer_t* r) {
  if (!r) throw InvalidError("null renderer handle");
  Renderer* s = reinterpret_cast<Renderer*>(r);
  if (g_multi_device) VNR_CUDA(cudaSetDevice(s->vol->device));
  return s;
}

VNR_EXPORT int vnr_renderer_create(vnr_volume_t* vh, vnr_renderer_t** out) {
  return guard([&] {
    Volume* v = V(vh);
    if (!out) throw InvalidError("null argument");
    *out = reinterpret_cast<vnr_renderer_t*>(new Renderer(v));
  });
}
VNR_EXPORT void vnr_renderer_release(vnr_renderer_t* r) { delete reinterpret_cast<Renderer*>(r); }
VNR_EXPORT int vnr_renderer_set_size(vnr_renderer_t* r, int w, int h) { return guard([&] { R(r)->resize(w, h); }); }
VNR_EXPORT int vnr_renderer_set_camera(vnr_renderer_t* r, const float* from, const float* at, const float* up, float fovy) {
  return guard([&] {
    Renderer* s = R(r);
    if (!from || !at || !up) throw InvalidError("null camera vector");
    for (int k = 0; k < 3; ++k) { s->cam_from[k] = from[k]; s->cam_at[k] = at[k]; s->cam_up[k] = up[k]; }
    if (fovy > 0.f) s->fovy = fovy;
    s->reset = true;                                   // renderer.h set_camera -> reset_frame
  });
}
VNR_EXPORT int vnr_renderer_set_mode(vnr_renderer_t* r, int mode) {
  return guard([&] {
    if (mode < 0 || mode >= 16) throw InvalidError("unknown rendering mode");   // api.h:85-86
    R(r)->mode = mode; R(r)->reset = true;
  });
}
VNR_EXPORT int vnr_renderer_set_groundtruth_source(vnr_renderer_t* r, int on) { return guard([&] { R(r)->gt_source = on != 0; R(r)->reset = true; }); }
VNR_EXPORT int vnr_renderer_set_sampling_rate(vnr_renderer_t* r, float rate) {
  return guard([&] { if (!(rate > 0.f)) throw InvalidError("sampling rate must be positive"); R(r)->sampling_rate = rate; R(r)->reset = true; });
}
VNR_EXPORT int vnr_renderer_set_density_scale(vnr_renderer_t* r, float s) { return guard([&] { R(r)->density_scale = s; R(r)->reset = true; }); }
// vnrVolumeSetClippingBox (api.h:146) -> MainRenderer::set_scene_clipbox (api.cpp:454): object-space box in [0,1]^3
VNR_EXPORT int vnr_renderer_set_clipping_box(vnr_renderer_t* r, const float* lower, const float* upper) {
  return guard([&] {
    Renderer* s = R(r);
    if (!lower || !upper) throw InvalidError("null clipping box");
    for (int k = 0; k < 3; ++k) { if (!(upper[k] > lower[k])) throw InvalidError("empty clipping box"); s->clip_lo[k] = lower[k]; s->clip_hi[k] = upper[k]; }
    s->reset = true;
  });
}
// vnrVolumeSetScaling (api.h:147, api.cpp:350-361): data transform = scale(s) * translate(-dims/2) * scale(dims)
VNR_EXPORT int vnr_renderer_set_scaling(vnr_renderer_t* r, const float* scale3) {
  return guard([&] {
    Renderer* s = R(r);
    if (!scale3) throw InvalidError("null scaling");
    for (int k = 0; k < 3; ++k) { if (!(scale3[k] > 0.f)) throw InvalidError("scaling must be positive"); s->scale[k] = scale3[k]; }
    s->reset = true;
  });
}
VNR_EXPORT int vnr_renderer_reset_accumulation(vnr_renderer_t* r) { return guard([&] { R(r)->reset = true; }); }
VNR_EXPORT int vnr_renderer_set_partition(vnr_renderer_t* r, int rank, int world) {
  return guard([&] {
    if (world < 1 || rank < 0 || rank >= world) throw InvalidError("bad partition");
    R(r)->part_rank = rank; R(r)->part_world = world; R(r)->reset = true;
  });
}
VNR_EXPORT int vnr_renderer_set_jitter_mode(vnr_renderer_t* r, int mode) {
  return guard([&] { if (mode != 0 && mode != 1) throw InvalidError("bad jitter mode"); R(r)->jitter_mode = mode; R(r)->reset = true; });
}
// measurement / test tap: ray order (8 x 4 pixel tiles per warp vs scanline) and sample-slot layout (depth-major per warp vs
// contiguous per ray) of the wavefront.  Frames do not depend on either.
VNR_EXPORT int vnr_renderer_set_layout(vnr_renderer_t* r, int tiled, int transpose) {
  return guard([&] { R(r)->tiled = tiled != 0; R(r)->transpose = transpose != 0; R(r)->reset = true; });
}
VNR_EXPORT int vnr_render(vnr_renderer_t* r) { return guard([&] { R(r)->render(); }); }
VNR_EXPORT const float* vnr_map_frame(vnr_renderer_t* r) {
  const float* p = nullptr;
  guard([&] { p = R(r)->map_frame(); });
  return p;
}
VNR_EXPORT int vnr_renderer_device_frame(vnr_renderer_t* r, void** d_rgba, void* stream_out) {
  return guard([&] {
    Renderer* s = R(r);
    if (!d_rgba) throw InvalidError("null argument");
    *d_rgba = s->last().frame.p;
    if (stream_out) *reinterpret_cast<cudaStream_t*>(stream_out) = s->last().stream;
  });
}
VNR_EXPORT int vnr_renderer_stats(vnr_renderer_t* r, uint64_t* s4) { return guard([&] { if (!s4) throw InvalidError("null argument"); R(r)->stats(s4); }); }

// decode entries emitted per wavefront round of the last frame (pass 0): out[r], r < max_rounds; *n_rounds = rounds with samples
VNR_EXPORT int vnr_renderer_round_counts(vnr_renderer_t* r, uint32_t* out, int max_rounds, int* n_rounds) {
  return guard([&] {
    Renderer* s = R(r);
    if (!out || max_rounds < 0) throw InvalidError("null argument");
    VNR_CUDA(cudaStreamSynchronize(s->last().stream));
    int used = 0;
    for (int k = 0; k < max_rounds; ++k) { out[k] = k < kMaxRounds ? s->last().h_counters[2 + k] : 0u; if (out[k]) used = k + 1; }
    if (n_rounds) *n_rounds = used;
  });
}
VNR_EXPORT int vnr_renderer_set_download(vnr_renderer_t* r, int on) { return guard([&] { R(r)->download = on != 0; }); }
// on (default): with the download enabled, the compositing kernels store finished pixels straight into the pinned host
// frame vnr_map_frame returns and the device frame buffer is left untouched; off: device frame + D2H copy after the frame
VNR_EXPORT int vnr_renderer_set_zero_copy(vnr_renderer_t* r, int on) { return guard([&] { R(r)->zero_copy = on != 0; R(r)->reset = true; }); }
VNR_EXPORT int vnr_renderer_download(vnr_renderer_t* r) { return guard([&] { R(r)->download_now(); }); }
VNR_EXPORT int vnr_renderer_set_profiling(vnr_renderer_t* r, int on) { return guard([&] { R(r)->profiling = on != 0; }); }
VNR_EXPORT int vnr_renderer_profile(vnr_renderer_t* r, float* decode_ms, int* decode_launches, uint64_t* kernel_launches) {
  return guard([&] { Renderer* s = R(r); s->profile(decode_ms, decode_launches); if (kernel_launches) *kernel_launches = s->last().launches; });
}
// the stream of the most recent frame's slot (with one frame in flight: THE stream of the renderer)
VNR_EXPORT int vnr_renderer_stream(vnr_renderer_t* r, void** stream) { return guard([&] { if (!stream) throw InvalidError("null argument"); *stream = (void*)R(r)->last().stream; }); }
// frame ring (MainRenderer's double buffer, renderer.h:84-94, generalised): n slots, each with its own stream and buffers
VNR_EXPORT int vnr_renderer_set_frames_in_flight(vnr_renderer_t* r, int n) { return guard([&] { R(r)->set_frames_in_flight(n); }); }
VNR_EXPORT int vnr_renderer_streams(vnr_renderer_t* r, void** streams, int max_streams, int* n_streams) {
  return guard([&] {
    Renderer* s = R(r);
    if (n_streams) *n_streams = (int)s->slots.size();
    for (int k = 0; streams && k < max_streams && k < (int)s->slots.size(); ++k) streams[k] = (void*)s->slot(k).stream;
  });
}
VNR_EXPORT int vnr_renderer_set_n_iters(vnr_renderer_t* r, int n) {
  return guard([&] { if (n < 1 || n > 32) throw InvalidError("n_iters must be in [1,32]"); R(r)->n_iters = n; R(r)->reset = true; });
}

// device-driven wavefront loop (CUDA graph WHILE node) on/off; off = bounded host-enqueued rounds
VNR_EXPORT int vnr_renderer_set_graph(vnr_renderer_t* r, int on) { return guard([&] { R(r)->use_graph = on != 0; }); }

// ---- multi-GPU frame gather over peer memory (no reference counterpart) ---------------------------
// Finished pixels of this renderer's partition are stored to `d_rgba` (float4[w*h]) instead of its own
// frame buffer: pass rank 0's frame buffer opened with vnr_ipc_open so that the compositing kernel writes
// straight into rank 0's memory over NVLink.  NULL restores the local buffer.
VNR_EXPORT int vnr_renderer_set_frame_target(vnr_renderer_t* r, void* d_rgba) {
  return guard([&] {
    Renderer* s = R(r);
    s->sync_all();
    if (s->slots.size() != 1) throw StateError("vnr_renderer_set_frame_target addresses a renderer with one frame in flight (multi-frame rings: vnr_comm)");
    s->slot(0).frame_target = reinterpret_cast<float4*>(d_rgba);
    s->reset = true;
  });
}
VNR_EXPORT int vnr_ipc_export(void* d_ptr, void* handle64) {
  return guard([&] {
    if (!d_ptr || !handle64) throw InvalidError("null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    VNR_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), d_ptr));
  });
}
VNR_EXPORT int vnr_ipc_open(const void* handle64, void** d_ptr) {
  return guard([&] {
    if (!d_ptr || !handle64) throw InvalidError("null argument");
    cudaIpcMemHandle_t h; memcpy(&h, handle64, sizeof h);
    VNR_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  });
}
VNR_EXPORT int vnr_ipc_close(void* d_ptr) { return guard([&] { if (d_ptr) VNR_CUDA(cudaIpcCloseMemHandle(d_ptr)); }); }

// ---- cross-rank barrier over peer memory (no reference counterpart; comm.cu) ---------------------------
// A stream-ordered barrier between the ranks of one NVSwitch box without a collective library call: every rank
// owns a small flag array mapped by all peers; `sync` launches ONE kernel of `world` threads on the caller's
// stream: thread t publishes this rank's epoch into peer t's array (system-scope release) and waits until peer
// t's epoch has arrived in the local array (acquire).  ~5 us instead of ~25 us for a 4-byte NCCL all-reduce.
// A peer that never arrives trips a 5 s timeout (error flag, no hang).
VNR_EXPORT int vnr_peer_barrier_create(void** out, void* handle64) {
  return guard([&] {
    if (!out || !handle64) throw InvalidError("null argument");
    require_device();
    std::unique_ptr<PeerBarrier> b(peer_barrier_create());
    VNR_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), b->local));
    *out = b.release();
  });
}
// all_handles: world x 64 bytes, rank-major
VNR_EXPORT int vnr_peer_barrier_attach(void* bh, int rank, int world, const void* all_handles) {
  return guard([&] {
    PeerBarrier* b = reinterpret_cast<PeerBarrier*>(bh);
    if (!b) throw InvalidError("null barrier");
    peer_barrier_attach_ipc(b, rank, world, all_handles);
  });
}
VNR_EXPORT int vnr_peer_barrier_sync(void* bh, void* stream) {
  return guard([&] {
    PeerBarrier* b = reinterpret_cast<PeerBarrier*>(bh);
    if (!b) throw InvalidError("null barrier");
    peer_barrier_sync(b, (cudaStream_t)stream);
  });
}
// epoch of the last barrier call that timed out (0 = healthy).  Reads a pinned host word the barrier kernel writes: it covers
// every barrier whose kernel has completed, so call it after a synchronisation you do anyway (frame mapped, loss read)
VNR_EXPORT int vnr_peer_barrier_check(void* bh, uint64_t* timed_out_epoch) {
  return guard([&] {
    PeerBarrier* b = reinterpret_cast<PeerBarrier*>(bh);
    if (!b || !timed_out_epoch) throw InvalidError("null argument");
    *timed_out_epoch = peer_barrier_timed_out(b);
  });
}
VNR_EXPORT void vnr_peer_barrier_release(void* bh) { delete reinterpret_cast<PeerBarrier*>(bh); }

// ---- communicators (comm.cu): multi-GPU behind the same calls -------------------------------------------------
static Comm* CM(vnr_comm_t* c) { if (!c) throw InvalidError("null communicator"); return reinterpret_cast<Comm*>(c); }
VNR_EXPORT int vnr_comm_init(int n_devices, vnr_comm_t** comms_out) {
  return guard([&] {
    if (!comms_out) throw InvalidError("null argument");
    require_device();
    std::vector<Comm*> cs = comm_create_local(n_devices);
    for (size_t k = 0; k < cs.size(); ++k) comms_out[k] = reinterpret_cast<vnr_comm_t*>(cs[k]);
    if (n_devices > 1) g_multi_device = true;
  });
}
VNR_EXPORT int vnr_comm_init_rank(int rank, int world, const char* rendezvous_name, vnr_comm_t** out) {
  return guard([&] {
    if (!out) throw InvalidError("null argument");
    require_device();
    *out = reinterpret_cast<vnr_comm_t*>(comm_create_rank(rank, world, rendezvous_name));
  });
}
VNR_EXPORT void vnr_comm_release(vnr_comm_t* c) { delete reinterpret_cast<Comm*>(c); }
VNR_EXPORT int vnr_comm_info(vnr_comm_t* c, int* rank, int* world, int* device) {
  return guard([&] { Comm* m = CM(c); if (rank) *rank = m->rank; if (world) *world = m->world; if (device) *device = m->device; });
}
VNR_EXPORT int vnr_comm_set_device(vnr_comm_t* c) { return guard([&] { VNR_CUDA(cudaSetDevice(CM(c)->device)); }); }
VNR_EXPORT int vnr_comm_barrier(vnr_comm_t* c) { return guard([&] { CM(c)->host_barrier(); }); }
VNR_EXPORT int vnr_volume_attach_comm(vnr_volume_t* vh, vnr_comm_t* c) { return guard([&] { comm_attach_volume(V(vh), CM(c)); }); }
VNR_EXPORT int vnr_volume_detach_comm(vnr_volume_t* vh) { return guard([&] { comm_detach_volume(V(vh)); }); }
VNR_EXPORT int vnr_renderer_attach_comm(vnr_renderer_t* r, vnr_comm_t* c) { return guard([&] { comm_attach_renderer(R(r), CM(c)); }); }
VNR_EXPORT int vnr_renderer_detach_comm(vnr_renderer_t* r) { return guard([&] { comm_detach_renderer(R(r)); }); }
