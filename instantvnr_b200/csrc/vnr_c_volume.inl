// vnr_c_volume.inl -- ground truth, macrocell, transfer function, training, params.json
#define VNR_TODO(name) { return guard([&] { throw UnsupportedError(name ": not implemented yet"); }); }
VNR_EXPORT int vnr_volume_load_params(vnr_volume_t*, const void*, size_t) VNR_TODO("vnr_volume_load_params")
VNR_EXPORT int vnr_volume_save_params(vnr_volume_t*, const void**, size_t*) VNR_TODO("vnr_volume_save_params")
VNR_EXPORT int vnr_params_peek(const void*, size_t, int*, int*, int*, const char**) VNR_TODO("vnr_params_peek")

VNR_EXPORT int vnr_volume_set_groundtruth_f32(vnr_volume_t* vh, const float* h_volume) {
  return guard([&] {
    Volume* v = V(vh);
    if (!h_volume) throw InvalidError("null volume data");
    const size_t n = (size_t)v->dims[0] * v->dims[1] * v->dims[2];
    v->gt.alloc(n);
    VNR_CUDA(cudaMemcpyAsync(v->gt.p, h_volume, n * sizeof(float), cudaMemcpyHostToDevice, v->stream));
    VNR_CUDA(cudaStreamSynchronize(v->stream));
    v->have_gt = true;
  });
}

VNR_EXPORT int vnr_volume_macrocell_from_groundtruth(vnr_volume_t* vh) {
  return guard([&] {
    Volume* v = V(vh);
    v->mc_range.zero(v->stream);
    macrocell_update_implicit(v, v->stream);
    macrocell_update_max_opacity(v, v->stream);
    v->mc_external = true;
    VNR_CUDA(cudaStreamSynchronize(v->stream));
  });
}

VNR_EXPORT int vnr_volume_get_macrocell(const vnr_volume_t* vh, int* mc_dims, float* h_value_range, float* h_max_opacity) {
  return guard([&] {
    const Volume* v = V(vh);
    if (mc_dims) for (int k = 0; k < 3; ++k) mc_dims[k] = v->mc_dims[k];
    VNR_CUDA(cudaStreamSynchronize(v->stream));
    if (h_value_range) VNR_CUDA(cudaMemcpy(h_value_range, v->mc_range.p, v->mc_range.bytes(), cudaMemcpyDeviceToHost));
    if (h_max_opacity) VNR_CUDA(cudaMemcpy(h_max_opacity, v->mc_maxop.p, v->mc_maxop.bytes(), cudaMemcpyDeviceToHost));
  });
}

VNR_EXPORT int vnr_volume_set_macrocell(vnr_volume_t* vh, const float* h_value_range) {
  return guard([&] {
    Volume* v = V(vh);
    if (!h_value_range) throw InvalidError("null argument");
    VNR_CUDA(cudaMemcpyAsync(v->mc_range.p, h_value_range, v->mc_range.bytes(), cudaMemcpyHostToDevice, v->stream));
    macrocell_update_max_opacity(v, v->stream);      // load_params_from_json :917
    VNR_CUDA(cudaStreamSynchronize(v->stream));
  });
}

VNR_EXPORT int vnr_volume_set_tfn(vnr_volume_t* vh, const float* rgb, int n_rgb, const float* alpha, int n_alpha, float lo, float hi) {
  return guard([&] {
    Volume* v = V(vh);
    if (n_rgb < 0 || n_alpha < 0 || (n_rgb && !rgb) || (n_alpha && !alpha)) throw InvalidError("bad transfer function arrays");
    if (n_alpha > 12288) throw InvalidError("transfer function too long");
    if (!(hi > lo)) throw InvalidError("empty transfer function value range");
    std::vector<float4> c(n_rgb);
    for (int i = 0; i < n_rgb; ++i) c[i] = make_float4(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2], 1.f);   // object.cpp:324-330
    VNR_CUDA(cudaStreamSynchronize(v->stream));
    v->tfn_color.alloc(n_rgb); v->tfn_alpha.alloc(n_alpha);
    if (n_rgb) VNR_CUDA(cudaMemcpyAsync(v->tfn_color.p, c.data(), n_rgb * sizeof(float4), cudaMemcpyHostToDevice, v->stream));
    if (n_alpha) VNR_CUDA(cudaMemcpyAsync(v->tfn_alpha.p, alpha, n_alpha * sizeof(float), cudaMemcpyHostToDevice, v->stream));
    v->n_color = n_rgb; v->n_alpha = n_alpha;
    // range clamped to the data range [0,1] (object.cpp:343-346)
    v->tfn_hi = std::min(1.f, hi); v->tfn_lo = std::max(0.f, lo);
    macrocell_update_max_opacity(v, v->stream);      // network.cu:749
    VNR_CUDA(cudaStreamSynchronize(v->stream));
  });
}

// NeuralVolume::train (network.cu:769-779): the macrocell is updated from every batch unless
// fast_mode with an external (ground-truth) macrocell; max opacity is refreshed when !fast_mode.
VNR_EXPORT int vnr_volume_train(vnr_volume_t* vh, int steps, int batch, int fast_mode, void* stream) {
  return guard([&] {
    Volume* v = V(vh);
    if (steps < 0 || batch < 0) throw InvalidError("negative steps / batch");
    cudaStream_t s = S(v, stream);
    const bool update_mc = !(fast_mode && v->mc_external);
    train_steps(v, steps, (size_t)batch, update_mc, s);
    if (!fast_mode) macrocell_update_max_opacity(v, s);
  });
}

VNR_EXPORT int vnr_volume_train_on(vnr_volume_t* vh, const float* d_xyz, const float* d_target, size_t n, void* stream) {
  return guard([&] {
    Volume* v = V(vh);
    if (!d_xyz || !d_target) throw InvalidError("null buffer");
    cudaStream_t s = S(v, stream);
    train_grads(v, d_xyz, d_target, n, n, s);
    optimizer_step(v, s);
  });
}

VNR_EXPORT int vnr_volume_train_grads(vnr_volume_t* vh, const float* d_xyz, const float* d_target, size_t n, size_t n_global, void* stream) {
  return guard([&] {
    Volume* v = V(vh);
    if (!d_xyz || !d_target) throw InvalidError("null buffer");
    if (n_global < n) throw InvalidError("n_global < n");
    train_grads(v, d_xyz, d_target, n, n_global, S(v, stream));
  });
}

VNR_EXPORT int vnr_volume_optimizer_step(vnr_volume_t* vh, void* stream) {
  return guard([&] { Volume* v = V(vh); optimizer_step(v, S(v, stream)); });
}

VNR_EXPORT int vnr_volume_grad_buffer(vnr_volume_t* vh, int which, void** d_grads, size_t* n_elems, int* is_f32) {
  return guard([&] {
    Volume* v = V(vh);
    train_ensure_buffers(v);
    if (which == 0) { if (d_grads) *d_grads = v->mlp_grads.p; if (n_elems) *n_elems = v->cfg.desc.n_mlp; if (is_f32) *is_f32 = 1; }
    else if (which == 1) { if (d_grads) *d_grads = v->grid_grads.p; if (n_elems) *n_elems = v->cfg.desc.n_grid; if (is_f32) *is_f32 = 0; }
    else throw InvalidError("gradient buffer index must be 0 (MLP, fp32) or 1 (grid, fp16)");
  });
}

VNR_EXPORT int vnr_volume_get_grads(vnr_volume_t* vh, float* h_mlp, uint16_t* h_grid) {
  return guard([&] {
    Volume* v = V(vh);
    train_ensure_buffers(v);
    VNR_CUDA(cudaStreamSynchronize(v->stream));
    if (h_mlp) VNR_CUDA(cudaMemcpy(h_mlp, v->mlp_grads.p, v->cfg.desc.n_mlp * sizeof(float), cudaMemcpyDeviceToHost));
    if (h_grid) VNR_CUDA(cudaMemcpy(h_grid, v->grid_grads.p, (size_t)v->cfg.desc.n_grid * sizeof(__half), cudaMemcpyDeviceToHost));
  });
}

VNR_EXPORT int vnr_volume_sample(vnr_volume_t* vh, float* d_xyz, float* d_target, size_t n, void* stream) {
  return guard([&] { Volume* v = V(vh); if (!d_xyz) throw InvalidError("null buffer"); sample_batch(v, d_xyz, d_target, n, S(v, stream)); });
}

VNR_EXPORT int vnr_volume_sampler_skip(vnr_volume_t* vh, uint64_t n_floats) {
  return guard([&] { V(vh)->sampler_rng.advance(n_floats); });
}

VNR_EXPORT int vnr_volume_sample_at(vnr_volume_t* vh, const float* h_xyz, float* h_out, size_t n, int hw_texture) {
  return guard([&] {
    Volume* v = V(vh);
    if (n && (!h_xyz || !h_out)) throw InvalidError("null buffer");
    DevBuf<float> x, y; x.alloc(3 * n); y.alloc(n);
    VNR_CUDA(cudaMemcpyAsync(x.p, h_xyz, 3 * n * sizeof(float), cudaMemcpyHostToDevice, v->stream));
    sample_at(v, x.p, y.p, n, hw_texture, v->stream);
    VNR_CUDA(cudaMemcpyAsync(h_out, y.p, n * sizeof(float), cudaMemcpyDeviceToHost, v->stream));
    VNR_CUDA(cudaStreamSynchronize(v->stream));
  });
}

// vnrNeuralVolumeGetTrainingStep / GetTrainingLoss: running mean of the per-step losses
// (tcnn_network.h:149-153)
VNR_EXPORT int vnr_volume_stats(vnr_volume_t* vh, uint64_t* step, double* loss) {
  return guard([&] {
    Volume* v = V(vh);
    double acc[2] = {0, 0};
    VNR_CUDA(cudaStreamSynchronize(v->stream));
    VNR_CUDA(cudaMemcpy(acc, v->loss_accum.p, sizeof acc, cudaMemcpyDeviceToHost));
    if (step) *step = v->train_step;
    if (loss) *loss = v->loss_count ? acc[0] / (double)v->loss_count : 0.0;
  });
}

VNR_EXPORT int vnr_volume_last_loss(vnr_volume_t* vh, double* loss) {
  return guard([&] {
    Volume* v = V(vh);
    double acc[2] = {0, 0};
    VNR_CUDA(cudaStreamSynchronize(v->stream));
    VNR_CUDA(cudaMemcpy(acc, v->loss_accum.p, sizeof acc, cudaMemcpyDeviceToHost));
    if (loss) *loss = acc[1];
  });
}
