// vnr_c_volume.inl -- ground truth, macrocell, transfer function, training, params.json
// ---- params.json (BSON) --------------------------------------------------------------------
// Layout written by NeuralVolume::save_params_to_json (core/network.cu:827-857) through
// nlohmann::json::to_bson: objects are std::map, i.e. keys in lexicographic order; unsigned and
// signed integers that fit int32 are BSON int32, floats are doubles, blobs are binary subtype 0.
//   { "macrocell": { "data": <vec2f[cells]>, "dims": {x,y,z}, "groundtruth": bool, "spacings": {x,y,z} },
//     "model": { "encoding": {...}, "loss": {...}, "network": {...} },
//     "parameters": { "n_params": N, "params_binary": <fp16[N]> },       (tcnn trainer.h:299-311)
//     "volume": { "dims": {x,y,z} } }
static mj::Value sorted_copy(const mj::Value& v) {
  if (v.type == mj::Value::ObjectT) {
    mj::Value o = mj::Value::make_object();
    std::vector<std::pair<std::string, mj::Value>> kv;
    for (auto& e : *v.obj) kv.emplace_back(e.first, sorted_copy(e.second));
    std::stable_sort(kv.begin(), kv.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    for (auto& e : kv) o.obj->push_back(e);
    return o;
  }
  if (v.type == mj::Value::ArrayT) {
    mj::Value a = mj::Value::make_array();
    for (auto& e : *v.arr) a.arr->push_back(sorted_copy(e));
    return a;
  }
  return v;
}

static mj::Value xyz_obj_int(int x, int y, int z) {
  mj::Value o = mj::Value::make_object();
  o.set("x", mj::Value::from_int(x)); o.set("y", mj::Value::from_int(y)); o.set("z", mj::Value::from_int(z));
  return o;
}

static int get_int(const mj::Value& o, const char* k) { return (int)o.at(k).num(); }

VNR_EXPORT int vnr_params_peek(const void* bson, size_t n, int* dx, int* dy, int* dz, const char** model_json) {
  static thread_local std::string model_text;
  return guard([&] {
    if (!bson || n < 5) throw InvalidError("empty params blob");
    mj::Value root;
    try { root = mj::bson::read(bson, n); } catch (const std::exception& e) { throw InvalidError(e.what()); }
    if (!root.contains("volume")) throw InvalidError("expecting a model config with volume dims tag");      // api.cpp:215
    const mj::Value& d = root.at("volume").at("dims");
    if (dx) *dx = get_int(d, "x");
    if (dy) *dy = get_int(d, "y");
    if (dz) *dz = get_int(d, "z");
    if (model_json) {
      model_text.clear();
      if (root.contains("model")) mj::dump(root.at("model"), model_text);
      *model_json = model_text.c_str();
    }
  });
}

// NeuralVolume::load_params_from_json (core/network.cu:879-939)
VNR_EXPORT int vnr_volume_load_params(vnr_volume_t* vh, const void* bson, size_t n) {
  return guard([&] {
    Volume* v = V(vh);
    if (!bson || n < 5) throw InvalidError("empty params blob");
    mj::Value root;
    try { root = mj::bson::read(bson, n); } catch (const std::exception& e) { throw InvalidError(e.what()); }
    if (root.contains("volume")) {
      const mj::Value& d = root.at("volume").at("dims");
      if (get_int(d, "x") != v->dims[0] || get_int(d, "y") != v->dims[1] || get_int(d, "z") != v->dims[2])
        throw InvalidError("mismatch data dimension");                                                   // network.cu:892
    }
    bool model_reset = false;
    if (root.contains("model")) {
      // deserialize_model (tcnn_network.h:163-221) ALWAYS rebuilds loss, optimizer, network and trainer from the stored config:
      // fresh Adam moments and per-parameter step counters, m_steps = 0, loss accumulators cleared -- also when the
      // architecture is unchanged.  The optimizer options are not part of the file ("optimizer" absent -> m_optimizer_opts).
      std::string text; mj::dump(root.at("model"), text);
      ModelConfig c = parse_model_config(text);
      const DecoderDesc &a = c.desc, &b = v->cfg.desc;
      const bool same = a.n_levels == b.n_levels && a.n_feat == b.n_feat && a.n_hidden == b.n_hidden && a.n_grid == b.n_grid && a.n_mlp == b.n_mlp &&
                        c.base_res == v->cfg.base_res && c.per_level_scale == v->cfg.per_level_scale;
      if (!same) {
        if (v->dp_world) throw StateError("detach the data-parallel peers before loading parameters of a different model");
        c.opt = v->cfg.opt;
        VNR_CUDA(cudaStreamSynchronize(v->stream));
        VNR_CUDA(cudaDeviceSynchronize());                  // renderers of this volume may still be decoding the old table
        v->cfg = c;
        v->params.alloc(c.n_params());
        v->master.release(); v->m1.release(); v->m2.release(); v->steps.release();
        v->grid_grads.release(); v->mlp_grads.release(); v->mlp_partial.release();
        v->have_params = false; v->grads_clean = false; v->decode_blob = 0;
      }
      v->have_opt = false;
      model_reset = true;
    }
    const mj::Value& P = root.contains("parameters") ? root.at("parameters") : root;                   // old format: params at the root
    if (!P.contains("params_binary") || P.at("params_binary").type != mj::Value::Binary) throw InvalidError("params blob has no params_binary");
    const std::string& blob = P.at("params_binary").s;
    if (blob.size() / 2 != v->cfg.n_params()) throw InvalidError("Can't set params because CPU buffer has the wrong size.");   // trainer.h:283
    if (root.contains("macrocell")) {
      const mj::Value& M = root.at("macrocell");
      const mj::Value& md = M.at("dims");
      const int mx = get_int(md, "x"), my = get_int(md, "y"), mz = get_int(md, "z");
      if (mx <= 0 || my <= 0 || mz <= 0) throw InvalidError("bad macrocell dims");
      if (mx != v->mc_dims[0] || my != v->mc_dims[1] || mz != v->mc_dims[2]) {
        // the reference re-allocates to the stored shape (network.cu:903-909); here the cell size is fixed at 16 voxels
        throw UnsupportedError("macrocell dims of the params file do not match ceil(dims/16)");
      }
      const mj::Value& data = M.at("data");
      if (data.type != mj::Value::Binary || data.s.size() != v->mc_range.bytes()) throw InvalidError("macrocell data has the wrong size");
      VNR_CUDA(cudaMemcpyAsync(v->mc_range.p, data.s.data(), data.s.size(), cudaMemcpyHostToDevice, v->stream));
      if (M.contains("groundtruth") && M.at("groundtruth").type == mj::Value::Bool) v->mc_external = M.at("groundtruth").b;
      macrocell_update_max_opacity(v, v->stream);                                                        // network.cu:912
      VNR_CUDA(cudaStreamSynchronize(v->stream));
    }
    std::vector<__half> h(blob.size() / 2);
    memcpy(h.data(), blob.data(), h.size() * 2);
    wait_for_frames(v, v->stream);
    VNR_CUDA(cudaStreamSynchronize(v->stream));            // in-flight training kernels still write params / master, frames read params
    VNR_CUDA(cudaMemcpy(v->params.p, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    upload_master_from_f16(v, h);
    v->have_params = true;
    if (!v->have_opt || model_reset) reset_optimizer_state(v);
  });
}

// NeuralVolume::save_params_to_json (core/network.cu:827-857) + json::to_bson (:865)
VNR_EXPORT int vnr_volume_save_params(vnr_volume_t* vh, const void** bson, size_t* n) {
  return guard([&] {
    Volume* v = V(vh);
    if (!bson || !n) throw InvalidError("null argument");
    if (!v->have_params) throw StateError("the neural volume has no parameters");
    VNR_CUDA(cudaStreamSynchronize(v->stream));
    mj::Value root = mj::Value::make_object();
    {
      mj::Value vol = mj::Value::make_object();
      vol.set("dims", xyz_obj_int(v->dims[0], v->dims[1], v->dims[2]));
      root.set("volume", vol);
    }
    {
      mj::Value mc = mj::Value::make_object();
      mc.set("groundtruth", mj::Value::from(v->mc_external));
      mc.set("dims", xyz_obj_int(v->mc_dims[0], v->mc_dims[1], v->mc_dims[2]));
      mj::Value sp = mj::Value::make_object();
      const char* ax[3] = {"x", "y", "z"};
      for (int k = 0; k < 3; ++k) sp.set(ax[k], mj::Value::from((double)(16.f / (float)v->dims[k])));   // macrocell.cu:200 (float -> double)
      mc.set("spacings", sp);
      std::vector<float> r(2 * v->cells());
      VNR_CUDA(cudaMemcpy(r.data(), v->mc_range.p, v->mc_range.bytes(), cudaMemcpyDeviceToHost));
      mc.set("data", mj::Value::binary(r.data(), r.size() * sizeof(float)));
      root.set("macrocell", mc);
    }
    {
      mj::Value p = mj::Value::make_object();
      const size_t np = v->cfg.n_params();
      p.set("n_params", mj::Value::from_int((int64_t)np));
      std::vector<__half> h(np);
      VNR_CUDA(cudaMemcpy(h.data(), v->params.p, np * 2, cudaMemcpyDeviceToHost));
      p.set("params_binary", mj::Value::binary(h.data(), np * 2));
      root.set("parameters", p);
    }
    root.set("model", mj::Parser::parse(v->cfg.model_json));
    v->blob = mj::bson::write(sorted_copy(root));
    *bson = v->blob.data();
    *n = v->blob.size();
  });
}

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

// same from a device buffer (device-to-device copy on the volume's stream): volumes that are produced or
// streamed on the GPU -- B200's 180 GB of HBM hold a 1024^3 float volume (4 GiB) outright, so the reference's
// out-of-core sampler (neural_sampler.cpp:488-1191) degenerates to in-core sampling
VNR_EXPORT int vnr_volume_set_groundtruth_device(vnr_volume_t* vh, const float* d_volume) {
  return guard([&] {
    Volume* v = V(vh);
    if (!d_volume) throw InvalidError("null volume data");
    const size_t n = (size_t)v->dims[0] * v->dims[1] * v->dims[2];
    v->gt.alloc(n);
    VNR_CUDA(cudaMemcpyAsync(v->gt.p, d_volume, n * sizeof(float), cudaMemcpyDeviceToDevice, v->stream));
    VNR_CUDA(cudaStreamSynchronize(v->stream));
    v->have_gt = true;
  });
}

// StaticSampler::load (core/samplers/neural_sampler.cpp:223-288) / OutOfCoreSampler (:488-1191): a raw structured
// volume file of any scalar type, streamed into HBM and normalised on the device (ingest.cu)
VNR_EXPORT int vnr_volume_set_groundtruth_file(vnr_volume_t* vh, const char* path, int value_type, uint64_t offset, int big_endian,
                                               float vmin, float vmax, float* range_out2) {
  return guard([&] {
    Volume* v = V(vh);
    if (!path) throw InvalidError("null path");
    load_groundtruth_file(v, path, value_type, offset, big_endian != 0, vmin, vmax, range_out2);
  });
}

// OutOfCoreSampler (core/samplers/neural_sampler.cpp:1040-1120): training draws from a pool of random slabs of the file
VNR_EXPORT int vnr_volume_set_groundtruth_outofcore(vnr_volume_t* vh, const char* path, int value_type, uint64_t offset, float vmin, float vmax,
                                                    uint32_t num_concurrent_blocks, uint32_t num_blocks) {
  return guard([&] {
    if (!path) throw InvalidError("null path");
    outofcore_open(V(vh), path, value_type, offset, vmin, vmax, num_concurrent_blocks, num_blocks);
  });
}
VNR_EXPORT int vnr_volume_outofcore_info(vnr_volume_t* vh, uint32_t* n_slots, uint32_t* n_refresh, uint64_t* slot_bytes, uint64_t* first_voxel,
                                         uint32_t* length, uint64_t* bytes_uploaded) {
  return guard([&] { outofcore_info(V(vh), n_slots, n_refresh, slot_bytes, first_voxel, length, bytes_uploaded); });
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
    wait_for_frames(v, v->stream);                       // frames in flight still classify with the current tables
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
    if (v->vcomm && v->vcomm->comm->world > 1) comm_train_steps(v, steps, (size_t)batch, update_mc, s);     // data parallel over the communicator
    else train_steps(v, steps, (size_t)batch, update_mc, s);
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
    if (v->vcomm && v->vcomm->resolved && v->vcomm->comm->world > 1) acc[0] = comm_global_loss(v, 0);     // the loss of the global batch: sum over ranks
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
    if (v->vcomm && v->vcomm->resolved && v->vcomm->comm->world > 1) acc[1] = comm_global_loss(v, 1);
    if (loss) *loss = acc[1];
  });
}

// ---- data-parallel hooks (no reference counterpart) ------------------------------------------------
VNR_EXPORT int vnr_volume_stream(vnr_volume_t* vh, void** stream) {
  return guard([&] { if (!stream) throw InvalidError("null argument"); *stream = (void*)V(vh)->stream; });
}
// MacroCell::update_explicit (core/macrocell.cu:42-73) on caller-provided samples
VNR_EXPORT int vnr_volume_macrocell_update(vnr_volume_t* vh, const float* d_xyz, const float* d_values, size_t n, void* stream) {
  return guard([&] {
    Volume* v = V(vh);
    if (n && (!d_xyz || !d_values)) throw InvalidError("null buffer");
    macrocell_update_explicit(v, d_xyz, d_values, n, S(v, stream));
  });
}
// device buffer of the value ranges: float[2*cells], (min - 1, max + 1) interleaved
VNR_EXPORT int vnr_volume_macrocell_buffer(vnr_volume_t* vh, void** d_range, size_t* n_floats) {
  return guard([&] {
    Volume* v = V(vh);
    if (d_range) *d_range = v->mc_range.p;
    if (n_floats) *n_floats = v->mc_range.n;
  });
}
// MacroCell::update_max_opacity (core/macrocell.cu:232-250) after the ranges changed
VNR_EXPORT int vnr_volume_macrocell_refresh(vnr_volume_t* vh, void* stream) {
  return guard([&] { Volume* v = V(vh); macrocell_update_max_opacity(v, S(v, stream)); });
}

// vnrNeuralVolumeGetPSNR (api.h:129; NeuralVolume::Impl::get_psnr core/network.cu:410-472)
VNR_EXPORT int vnr_volume_psnr(vnr_volume_t* vh, double* psnr) {
  return guard([&] { Volume* v = V(vh); if (!psnr) throw InvalidError("null argument"); *psnr = volume_psnr(v, v->stream); });
}
// vnrNeuralVolumeGetSSIM / GetTestingLoss / DecodeInference / DecodeReference (api.h:130-131,139-140; evaluate.cu)
VNR_EXPORT int vnr_volume_ssim(vnr_volume_t* vh, double* ssim, float* h_map) {
  return guard([&] { Volume* v = V(vh); if (!ssim) throw InvalidError("null argument"); *ssim = volume_ssim(v, h_map, v->stream); });
}
VNR_EXPORT int vnr_volume_test_loss(vnr_volume_t* vh, int batch, double* loss) {
  return guard([&] {
    Volume* v = V(vh);
    if (!loss) throw InvalidError("null argument");
    if (batch < 0) throw InvalidError("negative batch size");
    *loss = volume_test_loss(v, (size_t)batch, v->stream);
  });
}
VNR_EXPORT int vnr_volume_export(vnr_volume_t* vh, const char* path, int which, float* h_range2) {
  return guard([&] {
    Volume* v = V(vh);
    if (which != 0 && which != 1) throw InvalidError("which must be 0 (decoded volume) or 1 (ground truth)");
    volume_export(v, path, which, h_range2, v->stream);
  });
}

// ---- data-parallel optimizer over peer memory (train.cu) ---------------------------------------------
// handles192: three cudaIpcMemHandle_t (64 bytes each): parameters (fp16), hash-grid gradients (fp16), MLP gradients (fp32)
VNR_EXPORT int vnr_volume_dp_export(vnr_volume_t* vh, void* handles192) {
  return guard([&] {
    Volume* v = V(vh);
    if (!handles192) throw InvalidError("null argument");
    train_ensure_buffers(v);
    cudaIpcMemHandle_t* h = reinterpret_cast<cudaIpcMemHandle_t*>(handles192);
    VNR_CUDA(cudaIpcGetMemHandle(&h[0], v->params.p));
    VNR_CUDA(cudaIpcGetMemHandle(&h[1], v->grid_grads.p));
    VNR_CUDA(cudaIpcGetMemHandle(&h[2], v->mlp_grads.p));
  });
}
static void dp_detach_impl(Volume* v) {
  for (int r = 0; r < v->dp_world; ++r) {
    if (r == v->dp_rank) continue;
    if (v->dp_params[r]) cudaIpcCloseMemHandle(v->dp_params[r]);
    if (v->dp_grid_grads[r]) cudaIpcCloseMemHandle(v->dp_grid_grads[r]);
    if (v->dp_mlp_grads[r]) cudaIpcCloseMemHandle(v->dp_mlp_grads[r]);
  }
  for (int r = 0; r < kMaxPeers; ++r) v->dp_params[r] = v->dp_grid_grads[r] = v->dp_mlp_grads[r] = nullptr;
  v->dp_world = 0; v->dp_rank = 0;
}
// all_handles: world x 192 bytes, rank-major (what every rank's vnr_volume_dp_export wrote); maps the peers' buffers
VNR_EXPORT int vnr_volume_dp_attach(vnr_volume_t* vh, int rank, int world, const void* all_handles) {
  return guard([&] {
    Volume* v = V(vh);
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) throw InvalidError("bad data-parallel rank / world (at most 8 ranks)");
    if (world > 1 && !all_handles) throw InvalidError("null argument");
    train_ensure_buffers(v);
    train_preload_kernels(v);
    VNR_CUDA(cudaStreamSynchronize(v->stream));
    dp_detach_impl(v);
    v->dp_rank = rank; v->dp_world = world;
    for (int r = 0; r < world; ++r) {
      if (r == rank) { v->dp_params[r] = v->params.p; v->dp_grid_grads[r] = v->grid_grads.p; v->dp_mlp_grads[r] = v->mlp_grads.p; continue; }
      cudaIpcMemHandle_t h[3];
      memcpy(h, reinterpret_cast<const char*>(all_handles) + (size_t)r * sizeof h, sizeof h);
      VNR_CUDA(cudaIpcOpenMemHandle(&v->dp_params[r], h[0], cudaIpcMemLazyEnablePeerAccess));
      VNR_CUDA(cudaIpcOpenMemHandle(&v->dp_grid_grads[r], h[1], cudaIpcMemLazyEnablePeerAccess));
      VNR_CUDA(cudaIpcOpenMemHandle(&v->dp_mlp_grads[r], h[2], cudaIpcMemLazyEnablePeerAccess));
    }
    outofcore_set_rank(v, rank);                       // every rank refreshes its own random slabs
  });
}
VNR_EXPORT int vnr_volume_dp_detach(vnr_volume_t* vh) { return guard([&] { Volume* v = V(vh); VNR_CUDA(cudaStreamSynchronize(v->stream)); dp_detach_impl(v); }); }
// the fused reduce-scatter + Adam + all-gather; the caller places a cross-rank barrier before (all gradients complete)
// and after (all parameters complete), then calls vnr_volume_dp_finish_step (local gradient clear)
VNR_EXPORT int vnr_volume_dp_optimizer_step(vnr_volume_t* vh, void* stream) { return guard([&] { Volume* v = V(vh); dp_optimizer_step(v, S(v, stream)); }); }
VNR_EXPORT int vnr_volume_dp_finish_step(vnr_volume_t* vh, void* stream) { return guard([&] { Volume* v = V(vh); dp_finish_step(v, S(v, stream)); }); }

// vnrNeuralVolumeDecodeProgressive / GetNumberOfBlobs (api.h:134,137; core/network.cu:290-326)
VNR_EXPORT int vnr_volume_decode_progressive(vnr_volume_t* vh, void* stream) {
  return guard([&] { Volume* v = V(vh); decode_progressive(v, S(v, stream)); });
}
VNR_EXPORT int vnr_volume_num_blobs(const vnr_volume_t* vh, int* n) {
  return guard([&] { const Volume* v = V(vh); if (!n) throw InvalidError("null argument"); *n = (v->dims[2] + kSlicesPerBlob - 1) / kSlicesPerBlob; });
}
// test / export tap: the decoded volume (float[dx*dy*dz]); slices not decoded yet are zero
VNR_EXPORT int vnr_volume_get_decoded(vnr_volume_t* vh, float* h_out) {
  return guard([&] {
    Volume* v = V(vh);
    if (!h_out) throw InvalidError("null argument");
    if (!v->decoded.p) throw StateError("nothing has been decoded yet");
    VNR_CUDA(cudaStreamSynchronize(v->stream));
    VNR_CUDA(cudaMemcpy(h_out, v->decoded.p, v->decoded.bytes(), cudaMemcpyDeviceToHost));
  });
}

// ---- measurement taps of the training kernel ------------------------------------------------------
VNR_EXPORT int vnr_volume_train_debug(vnr_volume_t* vh, int variant, uint32_t flags, int profile) {
  return guard([&] {
    Volume* v = V(vh);
    if (variant < 0 || variant > 2) throw InvalidError("unknown training-kernel variant");
    VNR_CUDA(cudaStreamSynchronize(v->stream));
    v->train_variant = variant; v->train_flags = flags; v->train_prof_on = profile != 0;
  });
}
VNR_EXPORT int vnr_volume_train_profile(vnr_volume_t* vh, uint32_t* out, size_t max_words, int* n_ctas, int* words_per_cta) {
  return guard([&] {
    Volume* v = V(vh);
    const int wpc = train_profile_words();
    if (n_ctas) *n_ctas = (int)(v->train_prof.n / (size_t)wpc);
    if (words_per_cta) *words_per_cta = wpc;
    if (!out) return;
    VNR_CUDA(cudaStreamSynchronize(v->stream));
    const size_t n = std::min(max_words, v->train_prof.n);
    if (n) VNR_CUDA(cudaMemcpy(out, v->train_prof.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  });
}
