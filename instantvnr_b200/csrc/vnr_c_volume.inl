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

VNR_EXPORT int vnr_volume_train(vnr_volume_t*, int, int, int, void*) VNR_TODO("vnr_volume_train")
VNR_EXPORT int vnr_volume_train_on(vnr_volume_t*, const float*, const float*, size_t, void*) VNR_TODO("vnr_volume_train_on")
VNR_EXPORT int vnr_volume_train_grads(vnr_volume_t*, const float*, const float*, size_t, size_t, void*) VNR_TODO("vnr_volume_train_grads")
VNR_EXPORT int vnr_volume_optimizer_step(vnr_volume_t*, void*) VNR_TODO("vnr_volume_optimizer_step")
VNR_EXPORT int vnr_volume_grad_buffer(vnr_volume_t*, void**, size_t*, int*) VNR_TODO("vnr_volume_grad_buffer")
VNR_EXPORT int vnr_volume_sample(vnr_volume_t*, float*, float*, size_t, void*) VNR_TODO("vnr_volume_sample")
VNR_EXPORT int vnr_volume_sampler_skip(vnr_volume_t*, uint64_t) VNR_TODO("vnr_volume_sampler_skip")
VNR_EXPORT int vnr_volume_stats(vnr_volume_t*, uint64_t*, double*) VNR_TODO("vnr_volume_stats")
