/* vnr_c.h -- flat C ABI of the B200-native instantvnr hot path.
 *
 * This is the drop-in boundary: a C-ABI shared library (libvnr_b200.so) that sits under
 * the reference's C++ api.h surface.  Each entry point names the reference interface it
 * replaces (paths relative to the reference repo).  INTEGRATION.md shows the api.cpp-side
 * binding a maintainer would add.
 *
 * Conventions: every function returns 0 (VNR_OK) or a negative error code and never
 * throws across the boundary; vnr_last_error() returns the message of the last failure
 * on the calling thread.  Handles are opaque.  The caller owns every input buffer; output
 * buffers returned by the library (mapped frames, serialized blobs) are owned by the
 * handle and stay valid until the next call that produces the same kind of output or the
 * handle is released.  `stream` arguments are cudaStream_t passed as void* (NULL = the
 * handle's own stream).  Pointers named d_* are device pointers, h_* host pointers.
 * There is no CPU fallback: without a CUDA device every compute call fails with
 * VNR_ERR_CUDA.
 */
#ifndef VNR_C_H
#define VNR_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VNR_OK 0
#define VNR_ERR_INVALID (-1)     /* bad argument / bad config (api.cpp throws std::runtime_error) */
#define VNR_ERR_CUDA (-2)        /* CUDA runtime error (reference: CUDA_CHECK throws) */
#define VNR_ERR_UNSUPPORTED (-3) /* valid in the reference but outside this library's path */
#define VNR_ERR_STATE (-4)       /* call sequence error (e.g. train without ground truth) */

typedef struct vnr_volume vnr_volume_t;
typedef struct vnr_renderer vnr_renderer_t;

const char* vnr_last_error(void);
/* library / device probe: number of CUDA devices visible (0 => compute calls will fail) */
int vnr_device_count(void);

/* ---- neural volume ------------------------------------------------------------------ */

/* vnrCreateNeuralVolume(const vnrJson& config, vec3i dims)            api.h:123, api.cpp:190-204
 * model_json: text of the model config (example-model.json; // comments allowed). */
int vnr_volume_create(const char* model_json, int dx, int dy, int dz, vnr_volume_t** out);
void vnr_volume_release(vnr_volume_t* v);                              /* vnrRelease  api.h:185 */

/* sizes of the parameter blob (MLP matrices first, then grid levels;
 * tcnn network_with_input_encoding.h:134-151) */
int vnr_volume_model_info(const vnr_volume_t* v, uint64_t* n_params, uint64_t* n_mlp_params,
                          int* n_levels, int* n_features_per_level, int* n_hidden_layers);

/* tcnn Trainer::initialize_params (trainer.h:72-112): Xavier-uniform MLP + U(-1e-4,1e-4) grid
 * from pcg32 seeded through std::seed_seq{seed}.  The reference seeds with time(NULL)
 * (core/networks/tcnn_network.h:207); here the seed is explicit. Resets optimizer state. */
int vnr_volume_init_params(vnr_volume_t* v, uint32_t seed);

/* vnrNeuralVolumeSetModel (api.h:126; NeuralVolume::set_network_from_json(config) core/network.cu:731-741): replace the
 * network of an existing volume -- new encoding / MLP / optimizer from `model_json`, parameters initialised from `seed`,
 * training step back to 0 -- keeping dims, ground truth, sampler stream, macrocells and transfer function.  A config
 * that does not parse leaves the volume untouched. */
int vnr_volume_set_model(vnr_volume_t* v, const char* model_json, uint32_t seed);
/* tcnn Trainer::set_params / serialize: fp16 parameter blob           trainer.h:281-311 */
int vnr_volume_set_params_f16(vnr_volume_t* v, const uint16_t* h_params, size_t n);
int vnr_volume_get_params_f16(const vnr_volume_t* v, uint16_t* h_params, size_t n);

/* vnrNeuralVolumeSetParams / vnrCreateNeuralVolume(params) / vnrNeuralVolumeSerializeParams:
 * the reference's params.json BSON document (core/network.cu:827-939)   api.h:124-127,142-143 */
int vnr_volume_load_params(vnr_volume_t* v, const void* bson, size_t n);
int vnr_volume_save_params(vnr_volume_t* v, const void** bson, size_t* n);
/* dims / model stored in a params.json blob, to create the volume before loading it */
int vnr_params_peek(const void* bson, size_t n, int* dx, int* dy, int* dz, const char** model_json);

/* NeuralVolume::inference(len, d_input, d_output, stream)             core/network.h:102,
 * core/network.cu:1043-1052.  d_xyz: float[3*n] (x,y,z per sample, object space [0,1]^3),
 * d_out: float[n].  No padding requirement on n. */
int vnr_volume_decode(vnr_volume_t* v, const float* d_xyz, float* d_out, size_t n, void* stream);
/* same through host buffers (H2D + kernel + D2H + sync inside) */
int vnr_volume_decode_host(vnr_volume_t* v, const float* h_xyz, float* h_out, size_t n);
/* test tap: also returns the fp16 hash-grid features (row-major [n][enc_pad]) */
int vnr_volume_decode_debug(vnr_volume_t* v, const float* h_xyz, float* h_out, uint16_t* h_enc, size_t n);

/* measurement tap (no reference counterpart): the hash-grid gather alone at full occupancy, one folded
 * word per sample; bench.py times it to obtain the gather-rate ceiling of the decode roofline */
int vnr_volume_gather_probe(vnr_volume_t* v, const float* d_xyz, uint32_t* d_out, size_t n, void* stream);

/* measurement taps (no reference counterpart), independent of the product's gather code: one plain kernel at full
 * occupancy issuing n_ops random 16-byte operations over a fresh buffer of table_bytes (64 per thread, 8 in flight):
 * kind 0 = read-only loads (ld.global.nc.v4) -- a 46.7 MB buffer gives the L2 random-gather ceiling of the decode, a
 * 306.8 MB one the HBM random-gather ceiling (SURVEY 8d `l2_gather_gbs` / `hbm_gather_gbs`); kind 1 = fp16x8 vector
 * reductions (red.global.add.noftz.v4.f16x2), the ceiling of the hash-grid backward; kind 2 = a streaming copy of the
 * buffer; kind 3 = random 32-byte loads (one 256-bit load per sector); kinds 4-7 = one float4 frame of table_bytes
 * (1024 pixels wide) stored into PINNED HOST memory in scanline / 8x4-tile / 32-byte scanline / 16x2-tile order (what
 * bounds the zero-copy frame download).  Returns the fastest and the mean of `repeats` CUDA-event-timed launches (after one warm-up). */
int vnr_probe_memory(int kind, size_t table_bytes, size_t n_ops, int repeats, float* ms_best, float* ms_mean);
/* the same loads with the level structure of a hash-grid decode on uniform random coordinates: n_samples x 8 random 16-byte
 * loads per level, level l being level_entries[l] entries of 16 bytes (back to back) -- coarse levels stay cache-resident, levels
 * larger than the L2 miss in proportion.  The like-for-like gather ceiling for tables that exceed the L2 (no reference counterpart) */
int vnr_probe_levels(const uint32_t* level_entries, int n_levels, size_t n_samples, int repeats, float* ms_best, float* ms_mean);

/* measurement taps of the fused training kernel (train.cu): variant 1 = current MMA chain, 0 = the round-1 chain
 * (A/B); flags: 1 = the scatter groups issue no reductions, 2 = the gather groups issue no loads, 4 = the compute
 * group runs no MMA chain (results are then meaningless: timing only); profile != 0: the next training kernels record
 * per-CTA role timers, read back with vnr_volume_train_profile (words per CTA returned in *words_per_cta; out may be
 * NULL to query sizes; layout in train.cu) */
int vnr_volume_train_debug(vnr_volume_t* v, int variant, uint32_t flags, int profile);
int vnr_volume_train_profile(vnr_volume_t* v, uint32_t* out, size_t max_words, int* n_ctas, int* words_per_cta);

/* vnrCreateSimpleVolume + StaticSampler ground truth (core/samplers/neural_sampler.cu:86-128):
 * float32 volume of dims dx*dy*dz (x fastest), already normalised to [0,1]. */
int vnr_volume_set_groundtruth_f32(vnr_volume_t* v, const float* h_volume);
/* StaticSampler::load (core/samplers/neural_sampler.cpp:223-288) and the out-of-core samplers (:488-1191): a raw
 * structured volume file (x fastest) of scalar type `value_type` (the reference's ValueType, core/mathdef.h:51-65:
 * 0 uint8, 1 int8, 2 uint16, 3 int16, 4 uint32, 5 int32, 8 float, 12 double) starting at byte `offset`, streamed
 * into HBM through pinned double buffers and normalised on the device: clamp((v - vmin) / (vmax - vmin), 0, 1)
 * (convert_volume :176-210).  vmin >= vmax: the range is computed from the data (one extra pass over the file).
 * range_out2 (optional) receives the unnormalised range used.  The volume then lives in HBM; training samples it
 * in-core (a 1024^3 float volume is 4 GiB of the 180 GB). */
int vnr_volume_set_groundtruth_file(vnr_volume_t* v, const char* path, int value_type, uint64_t offset, int big_endian,
                                    float vmin, float vmax, float* range_out2);
/* OutOfCoreSampler (core/samplers/neural_sampler.cpp:1040-1120) with its RandomBuffer (:488-668): the volume stays in the
 * file; a pool of `num_blocks` random slabs (full x-rows x ceil(32 KiB / row bytes) y-rows x 1 z-slice, + 1 ghost row / slice
 * per side), kept in HBM in the file's scalar type (little endian), is refreshed by `num_concurrent_blocks` slabs per
 * training step and sampled on the device: random slab, random voxel, random point in its cell, trilinear interpolation
 * of the values normalised with [vmin, vmax] BEFORE interpolating.  0 for either count: environment VNR_NUM_CONCURRENT_BLOCKS
 * (1024) / VNR_NUM_BLOCKS (64 x concurrent), as the reference.  A valid range is required (:1068-1070).  Afterwards
 * vnr_volume_train / vnr_volume_sample draw from the pool (five uniforms of the sampler's pcg32 stream per sample).  The file is
 * mapped and registered with the CUDA driver when that is possible (the refresh is then pulled by the GPU out of the page cache,
 * no host copy); otherwise host threads read the slabs into pinned staging buffers. */
int vnr_volume_set_groundtruth_outofcore(vnr_volume_t* v, const char* path, int value_type, uint64_t offset, float vmin, float vmax,
                                         uint32_t num_concurrent_blocks, uint32_t num_blocks);
/* pool geometry and, for test restatement, the slab table the next sample call will use (first file voxel and voxel count
 * per slot; arrays of n_slots entries or NULL); bytes_uploaded counts host->device slab traffic so far */
int vnr_volume_outofcore_info(vnr_volume_t* v, uint32_t* n_slots, uint32_t* n_refresh, uint64_t* slot_bytes, uint64_t* first_voxel,
                              uint32_t* length, uint64_t* bytes_uploaded);
/* same from a device buffer (float[dx*dy*dz], copied): volumes produced / streamed on the GPU */
int vnr_volume_set_groundtruth_device(vnr_volume_t* v, const float* d_volume);
/* MacroCell::compute_everything (core/macrocell.cu:221-230): value ranges from ground truth */
int vnr_volume_macrocell_from_groundtruth(vnr_volume_t* v);
int vnr_volume_get_macrocell(const vnr_volume_t* v, int* mc_dims, float* h_value_range /*2*cells or NULL*/,
                             float* h_max_opacity /*cells or NULL*/);
int vnr_volume_set_macrocell(vnr_volume_t* v, const float* h_value_range /*2*cells, stored with -1/+1 offset*/);

/* NeuralVolume::set_transfer_function (core/network.cu:743-760) + MacroCell::update_max_opacity.
 * rgb: float[3*n_rgb]; alpha: float[n_alpha] (the .y of the reference's vec2f list). */
int vnr_volume_set_tfn(vnr_volume_t* v, const float* rgb, int n_rgb, const float* alpha, int n_alpha, float lo, float hi);

/* vnrNeuralVolumeTrain(v, steps, fast_mode)                            api.h:136, core/network.cu:769-779
 * batch: samples per step (reference hard-codes 1<<16, core/network.cu:183; 0 = that default). */
int vnr_volume_train(vnr_volume_t* v, int steps, int batch, int fast_mode, void* stream);
/* One training step on caller-provided samples (AbstractNetwork::train, tcnn_network.h:223-252).
 * d_xyz float[3*n], d_target float[n]; n must be a multiple of 128. */
int vnr_volume_train_on(vnr_volume_t* v, const float* d_xyz, const float* d_target, size_t n, void* stream);
/* Data-parallel split of a step: gradients only (fwd+loss+bwd), then the optimizer.  Between
 * the two the caller all-reduces the gradient buffer (vnr_volume_grad_buffer).  Calls before the next optimizer step
 * ACCUMULATE (hash-grid and MLP gradients alike): k calls on k batches with n_global = their total size equal one step
 * on the concatenated batch.
 * Arithmetic: the MLP weight gradients are accumulated in HALF like the reference's split-K GEMMs (tcnn cutlass_matmul.h:83;
 * fully_fused_mlp.cu:863-922): one rounding per 16 samples, one K-slice per SM, slices summed in half. */
int vnr_volume_train_grads(vnr_volume_t* v, const float* d_xyz, const float* d_target, size_t n, size_t n_global, void* stream);
int vnr_volume_optimizer_step(vnr_volume_t* v, void* stream);
/* which = 0: MLP weight gradients (fp32, n_mlp elements); which = 1: hash-grid gradients (fp16,
 * n_grid elements).  Both carry the loss scale (x128, tcnn trainer.h:234). */
int vnr_volume_grad_buffer(vnr_volume_t* v, int which, void** d_grads, size_t* n_elems, int* is_f32);
/* test tap: copy the current gradients to the host (either pointer may be NULL) */
int vnr_volume_get_grads(vnr_volume_t* v, float* h_mlp /*n_mlp*/, uint16_t* h_grid /*n_grid, fp16*/);
/* StaticSampler::sample (core/samplers/neural_sampler.cu:131-164) into device buffers */
int vnr_volume_sample(vnr_volume_t* v, float* d_xyz, float* d_target, size_t n, void* stream);
/* advance the sampler's pcg32 stream (rank r of a data-parallel job skips r*3*n per step) */
int vnr_volume_sampler_skip(vnr_volume_t* v, uint64_t n_floats);
/* test tap: trilinear ground-truth lookup at host coordinates.  hw_texture = 0: the product's
 * software filter (1.8 fixed-point weights); 1: a real CUDA 3-D texture as the reference uses
 * (tex3D<float>, linear filter, normalized coordinates, clamp). */
int vnr_volume_sample_at(vnr_volume_t* v, const float* h_xyz, float* h_out, size_t n, int hw_texture);

/* data-parallel hooks (no reference counterpart): the cudaStream_t the volume's own work is enqueued on;
 * MacroCell::update_explicit (core/macrocell.cu:42-73) on caller-provided samples; the value-range buffer
 * (float[2*cells] = (min-1, max+1) per cell, so a min / max all-reduce of the even / odd elements merges
 * ranks); MacroCell::update_max_opacity (:232-250) after the merge */
int vnr_volume_stream(vnr_volume_t* v, void** stream);
int vnr_volume_macrocell_update(vnr_volume_t* v, const float* d_xyz, const float* d_values, size_t n, void* stream);
int vnr_volume_macrocell_buffer(vnr_volume_t* v, void** d_range, size_t* n_floats);
int vnr_volume_macrocell_refresh(vnr_volume_t* v, void* stream);

/* Data-parallel optimizer over peer memory (no reference counterpart; one 8-GPU NVSwitch box): rank r owns an
 * interleaved 1/world of the hash-grid parameters; ONE kernel reads every rank's gradients over NVLink, runs Adam on
 * the owned shard and stores the new fp16 parameters into every rank's buffer (reduce-scatter + Adam + all-gather).
 * Sequence per step on every rank: vnr_volume_train_grads -> barrier -> vnr_volume_dp_optimizer_step -> barrier ->
 * vnr_volume_dp_finish_step.  export: 3 x 64-byte cudaIpcMemHandle_t (params, grid gradients, MLP gradients);
 * attach: all ranks' exports concatenated rank-major. */
int vnr_volume_dp_export(vnr_volume_t* v, void* handles192);
int vnr_volume_dp_attach(vnr_volume_t* v, int rank, int world, const void* all_handles);
int vnr_volume_dp_detach(vnr_volume_t* v);
int vnr_volume_dp_optimizer_step(vnr_volume_t* v, void* stream);
int vnr_volume_dp_finish_step(vnr_volume_t* v, void* stream);

/* vnrNeuralVolumeGetTrainingStep / GetTrainingLoss                     api.h:132-133 */
int vnr_volume_stats(vnr_volume_t* v, uint64_t* step, double* loss);
/* loss of the most recent step (sum over the batch of |y - t| / N) */
int vnr_volume_last_loss(vnr_volume_t* v, double* loss);

/* vnrNeuralVolumeGetPSNR (api.h:129; NeuralVolume::Impl::get_psnr core/network.cu:410-472): decode every voxel
 * centre, 10 log10(range^2 / mse) against the ground truth */
int vnr_volume_psnr(vnr_volume_t* v, double* psnr);
/* vnrNeuralVolumeGetSSIM (api.h:130; get_mssim core/network.cu:474-549, compute_ssim :70-125): mean structural
 * similarity of the decoded volume against the ground truth over 7^3 uniform windows (sample covariance, K1 0.01,
 * K2 0.03, data range 1), averaged over the (dims-6)^3 window origins.  h_map (optional, (dx-6)(dy-6)(dz-6) floats)
 * receives the per-window values (test tap). */
int vnr_volume_ssim(vnr_volume_t* v, double* ssim, float* h_map);
/* vnrNeuralVolumeGetTestingLoss (api.h:131; NeuralVolume::Impl::test core/network.cu:261-288): mean |decode - target|
 * over one fresh batch of the training sampler (batch 0 = 65536; the draw advances the sampler stream, as the
 * reference's shared static generator does) */
int vnr_volume_test_loss(vnr_volume_t* v, int batch, double* loss);
/* vnrNeuralVolumeDecodeInference / DecodeReference (api.h:139-140; save_inference_volume / save_reference_volume
 * core/network.cu:327-408): write the decoded volume (which = 0) or the ground truth (which = 1) as dz records of
 * next_multiple(dx*dy, 256) floats; h_range2 (optional) receives the min / max written */
int vnr_volume_export(vnr_volume_t* v, const char* path, int which, float* h_range2);

/* vnrNeuralVolumeDecodeProgressive (api.h:137; infer_progressively_decode_volume core/network.cu:290-326): decode the
 * next blob of 16 z-slices of voxel centres into the decoded volume that the "decoding" rendering modes march;
 * vnrNeuralVolumeGetNumberOfBlobs (api.h:134) calls cover the volume once; get_decoded copies it out (test tap). */
int vnr_volume_decode_progressive(vnr_volume_t* v, void* stream);
int vnr_volume_num_blobs(const vnr_volume_t* v, int* n);
int vnr_volume_get_decoded(vnr_volume_t* v, float* h_out);

/* ---- scene descriptions ---------------------------------------------------------------- */
/* The scene JSON the reference's apps pass to vnrCreateSimpleVolume / vnrCreateCamera / vnrCreateTransferFunction
 * (api.h:104-106,117,155; serializer.cpp:138-477): "version": "VIDI3D" (default; dataSource[] of
 * REGULAR_GRID_RAW_BINARY files, view.volume.scalarMappingRange[Unnormalized], view.camera) or "DIVA" (volume{dims,
 * type, range, filename[, bigendian]}).  is_path != 0: `json` names a file (a vnrJson that is_string(), api.cpp:75-80).
 * Host-only: no CUDA device is needed. */
typedef struct vnr_scene vnr_scene_t;
int vnr_scene_create(const char* json, int is_path, vnr_scene_t** out);
void vnr_scene_release(vnr_scene_t* s);
/* MultiVolume (core/instantvnr_types.h:40-56): dims, ValueType, number of time steps, unnormalised value range
 * (has_range = 0: compute it from the data, as an empty range1f makes StaticSampler::load do) */
int vnr_scene_volume(const vnr_scene_t* s, int* dims3, int* value_type, int* n_timesteps, float* range2, int* has_range);
/* MultiVolume::File of time step t; the string lives as long as the scene */
int vnr_scene_timestep(const vnr_scene_t* s, int t, const char** filename, uint64_t* offset, int* big_endian);
/* create_json_camera_stringify (serializer.cpp:395-407): eye / center shifted by -dims/2 into the centred world box.
 * VNR_ERR_STATE when the scene has no camera (DIVA scenes: "TODO" in the reference) */
int vnr_scene_camera(const vnr_scene_t* s, float* from3, float* at3, float* up3, float* fovy);
/* create_json_tfn_stringify (:371-377): value range of the transfer function (scalarMappingRangeUnnormalized, or
 * scalarMappingRange x the integer type's maximum); has_range = 0 when neither key exists.  The colour / alpha TABLE of
 * the reference comes from tfn::loadTransferFunction of the un-vendored OVR tfn module and is not restated: tables are
 * returned only for the explicit form {"colors": [[r,g,b],..], "alphas": [[pos,alpha],..]}; otherwise n_rgb = n_alpha = 0
 * and, when the scene carries some other transferFunction object, the call returns VNR_ERR_UNSUPPORTED after filling
 * the range.  Pointers live as long as the scene. */
int vnr_scene_tfn(const vnr_scene_t* s, const float** rgb, int* n_rgb, const float** alpha_pairs, int* n_alpha, float* range2, int* has_range);

/* ---- renderer ------------------------------------------------------------------------- */

int vnr_renderer_create(vnr_volume_t* v, vnr_renderer_t** out);        /* vnrCreateRenderer api.h:168 */
void vnr_renderer_release(vnr_renderer_t* r);
int vnr_renderer_set_size(vnr_renderer_t* r, int width, int height);   /* SetFramebufferSize :169 */
int vnr_renderer_set_camera(vnr_renderer_t* r, const float* from, const float* at, const float* up, float fovy); /* :171 */
int vnr_renderer_set_mode(vnr_renderer_t* r, int mode);                /* vnrRenderMode, api.h:36-60.  Ray marching: 4 / 7 / 10 march
                                                                          the progressively decoded volume, 5 / 8 / 11 and 6 / 9 / 12 decode
                                                                          the network per sample (no / gradient / single-shade shading);
                                                                          path tracing: 13 decoded volume, 14 / 15 network per collision.
                                                                          0-3 (OptiX): vnr_render returns VNR_ERR_UNSUPPORTED */
/* what a SimpleVolume renderer marches (vnrCreateRenderer on a simple volume, api.cpp:441-452): 1 = the ground-truth
 * volume of `v` through the same wavefront (the "GT render" frames are compared against); 0 = per mode */
int vnr_renderer_set_groundtruth_source(vnr_renderer_t* r, int on);
int vnr_renderer_set_sampling_rate(vnr_renderer_t* r, float rate);     /* :174 */
int vnr_renderer_set_density_scale(vnr_renderer_t* r, float scale);    /* :175 */
/* vnrVolumeSetScaling (api.h:147; api.cpp:350-361): per-axis scale of the world box, multiplied onto the default
 * data transform translate(-dims/2) * scale(dims) */
int vnr_renderer_set_scaling(vnr_renderer_t* r, const float* scale3);
int vnr_renderer_reset_accumulation(vnr_renderer_t* r);                /* :176 */
/* vnrVolumeSetClippingBox (api.h:146, applied by vnrCreateRenderer api.cpp:454): object-space box inside [0,1]^3 */
int vnr_renderer_set_clipping_box(vnr_renderer_t* r, const float* lower3, const float* upper3);
/* pixel subset for tile-parallel multi-GPU rendering: this renderer handles pixel tiles
 * (64x... row blocks) with index % world == rank.  Default (0,1) = whole frame. */
int vnr_renderer_set_partition(vnr_renderer_t* r, int rank, int world);
/* jitter source: 0 = gdt::LCG<16>(frame_index, pixel) as the reference, 1 = fixed 0.5 */
int vnr_renderer_set_jitter_mode(vnr_renderer_t* r, int mode);
/* no reference counterpart (measurement / test tap): how the wavefront deals rays to warps (1: 8 x 4 pixel tiles, 0: scanline
 * segments) and lays out the sample slots of a warp (1: depth-major, 0: contiguous per ray).  Defaults 1, 1; frames are
 * identical under all four combinations. */
int vnr_renderer_set_layout(vnr_renderer_t* r, int tiled, int transpose);
int vnr_render(vnr_renderer_t* r);                                     /* vnrRender :177 (async) */
/* vnrRendererMapFrame :178: syncs, returns host float4[w*h] valid until the second-next map */
const float* vnr_map_frame(vnr_renderer_t* r);
/* device frame buffer (float4[w*h]) of the last vnr_render, for on-device gathers */
int vnr_renderer_device_frame(vnr_renderer_t* r, void** d_rgba, void* stream_out);
/* counters of the last frame: [0] rays that hit the volume, [1] samples decoded,
 * [2] samples composited, [3] wavefront rounds */
int vnr_renderer_stats(vnr_renderer_t* r, uint64_t* stats4);

/* measurement tap: decode entries emitted per wavefront round of the last frame (the reference reads the same counter back
 * every round, method_raymarching.cu:923-929) */
int vnr_renderer_round_counts(vnr_renderer_t* r, uint32_t* out, int max_rounds, int* n_rounds);

/* MainRenderer::framebuffer_skip_download (renderer.cpp:132): 0 = keep frames on the device */
int vnr_renderer_set_download(vnr_renderer_t* r, int on);
/* Zero-copy download (default on; no reference counterpart -- the reference copies the frame after it is complete,
 * renderer.cpp:133): while the download is enabled and no frame target is set, finished pixels are stored by the
 * compositing kernels straight into the pinned host frame vnr_map_frame returns, overlapping PCIe with the rest of
 * the wavefront; the device frame buffer (vnr_renderer_device_frame) is then NOT updated.  0 = device frame + one
 * device->host copy after the frame. */
int vnr_renderer_set_zero_copy(vnr_renderer_t* r, int on);
/* framebuffer.download_async (framebuffer.h:35) on demand, for callers that disabled the automatic one:
 * enqueues the device->host copy of the current frame; vnr_map_frame then waits for it */
int vnr_renderer_download(vnr_renderer_t* r);
/* measurement taps: CUDA events around every decode launch of the next frames; summed device time of
 * the non-empty decode launches of the last frame, their number, and all kernels launched by it */
int vnr_renderer_set_profiling(vnr_renderer_t* r, int on);
int vnr_renderer_profile(vnr_renderer_t* r, float* decode_ms, int* decode_launches, uint64_t* kernel_launches);
/* the cudaStream_t all work of this renderer is enqueued on (several frames in flight: the stream of the most recent frame) */
int vnr_renderer_stream(vnr_renderer_t* r, void** stream);
/* Frames in flight (MainRenderer's double-buffered framebuffer, renderer.h:84-94, framebuffer.h:73-77, generalised): the
 * renderer owns a ring of n (1..8, default 1, env VNR_FRAMES_IN_FLIGHT) frame slots, each with its own stream, ray /
 * sample / value buffers, accumulation buffer, captured wavefront graph and pinned host frames.  vnr_render enqueues the
 * frame on the next slot and returns, so consecutive frames overlap on the device; vnr_map_frame returns the OLDEST rendered
 * frame that has not been mapped (n = 1: the frame just rendered, as the reference).  A frame that accumulates onto the
 * previous one (frame_index > 1) waits for it on the device.  Rendering into a full ring drops the oldest unmapped frame. */
int vnr_renderer_set_frames_in_flight(vnr_renderer_t* r, int n);
/* the slots' streams (to bracket a run of frames with events); n_streams receives the ring depth */
int vnr_renderer_streams(vnr_renderer_t* r, void** streams, int max_streams, int* n_streams);
/* samples per ray per wavefront round (N_ITERS, env VNR_RM_N_ITERS; method_raymarching.cu:30-40): 1..32, default 16 as the
 * reference.  Frames do not depend on it (a ray's samples are the same, only the round that decodes them changes); unshaded
 * marching (modes 4-6) honours values above 16 -- fewer, fuller rounds -- the shaded passes cap it at 16. */
int vnr_renderer_set_n_iters(vnr_renderer_t* r, int n);

/* device-driven wavefront loop (default on): the per-round host round trip of iterative_ray_compaction
 * (method_raymarching.cu:923-929) is replaced by a CUDA-graph WHILE node; 0 = bounded host-enqueued rounds */
int vnr_renderer_set_graph(vnr_renderer_t* r, int on);

/* multi-GPU frame gather over peer memory (no reference counterpart): finished pixels of this renderer's
 * partition are stored to d_rgba (float4[w*h], normally rank 0's frame buffer opened with vnr_ipc_open)
 * by the compositing kernel itself; NULL restores the local frame buffer. */
int vnr_renderer_set_frame_target(vnr_renderer_t* r, void* d_rgba);
/* cudaIpcMemHandle_t (64 bytes) of a device allocation of this process / mapping of a peer's allocation */
int vnr_ipc_export(void* d_ptr, void* handle64);
int vnr_ipc_open(const void* handle64, void** d_ptr);
int vnr_ipc_close(void* d_ptr);

/* Stream-ordered barrier between the ranks of one NVSwitch box over peer memory (no reference counterpart): one
 * `world`-thread kernel publishes this rank's epoch into every peer's flag array and waits for theirs (system-scope
 * release / acquire); replaces a 4-byte NCCL all-reduce (~25 us) at ~5 us.  create: returns the 64-byte IPC handle
 * of the local flags; attach: all ranks' handles rank-major; a peer that never arrives trips a 5 s timeout that
 * vnr_peer_barrier_check reports (it reads a pinned host word: call it after a synchronisation that covers the barrier,
 * e.g. once the frame is mapped or the loss read).  Volumes / renderers attached to a communicator check their own
 * barriers there and fail with VNR_ERR_STATE (vnr_map_frame, vnr_volume_stats, vnr_volume_last_loss). */
int vnr_peer_barrier_create(void** barrier, void* handle64);
int vnr_peer_barrier_attach(void* barrier, int rank, int world, const void* all_handles);
int vnr_peer_barrier_sync(void* barrier, void* stream);
int vnr_peer_barrier_check(void* barrier, uint64_t* timed_out_epoch);
void vnr_peer_barrier_release(void* barrier);

/* ---- communicators: multi-GPU behind the same calls (SURVEY 8b `vnr_comm_init(n_devices)`; no reference counterpart, the
 * reference runs on one GPU) -------------------------------------------------------------------------------------------
 * One NVSwitch box, at most 8 GPUs.  Two ways to span them:
 *   vnr_comm_init(n, comms[n])            one process, n devices: comms[r] belongs to device r.  Create the objects of
 *                                         rank r with that device current (vnr_comm_set_device(comms[r])); afterwards every
 *                                         entry point makes the object's device current itself.
 *   vnr_comm_init_rank(rank, world, name) one process per GPU (torchrun, mpirun): the ranks meet in a POSIX shared-memory
 *                                         segment called `name` (unique per job, e.g. "vnr-<port>-<launcher pid>"); the device
 *                                         is the current one.
 * Attaching is collective (every rank attaches its k-th volume / renderer; in the one-process form the exchange happens when
 * the last rank attaches):
 *   vnr_volume_attach_comm    rank 0's parameters, macrocell value ranges and sampler stream are replicated, the optimizer
 *                             restarts (as a new Trainer), and vnr_volume_train(v, steps, batch, ..) becomes synchronous data
 *                             parallel: `batch` samples PER RANK per step (rank r takes the r-th of `world` consecutive batches of
 *                             the one sampler stream), loss normalised by the global batch, then reduce-scatter + Adam +
 *                             all-gather in one kernel over NVLink peer memory; value ranges are merged at the end of the call.
 *                             The result equals one process accumulating `world` consecutive batches per optimizer step.
 *                             vnr_volume_stats / vnr_volume_last_loss report the loss of the global batch.
 *   vnr_renderer_attach_comm  after vnr_renderer_set_size (+ frames in flight): rank r renders the image strips r, r + world, ..;
 *                             finished pixels are stored by the compositing kernels into pinned host frames shared by all ranks
 *                             (each GPU delivers its strips over its own PCIe link) or, with the download disabled, into rank
 *                             0's device frame over NVLink; vnr_render on every rank, vnr_map_frame on rank 0.  Camera, mode,
 *                             transfer function, sampling rate are set identically on every rank by the caller.
 * Ordering between ranks is a stream-ordered peer barrier kernel; nothing goes through the host or a collective library. */
typedef struct vnr_comm vnr_comm_t;
int vnr_comm_init(int n_devices, vnr_comm_t** comms_out /* n_devices entries */);
int vnr_comm_init_rank(int rank, int world, const char* rendezvous_name, vnr_comm_t** out);
void vnr_comm_release(vnr_comm_t* c);
int vnr_comm_info(vnr_comm_t* c, int* rank, int* world, int* device);
int vnr_comm_set_device(vnr_comm_t* c);                 /* cudaSetDevice(the communicator's device) */
int vnr_comm_barrier(vnr_comm_t* c);                    /* host barrier between the ranks (one process per GPU) */
int vnr_volume_attach_comm(vnr_volume_t* v, vnr_comm_t* c);
int vnr_volume_detach_comm(vnr_volume_t* v);
int vnr_renderer_attach_comm(vnr_renderer_t* r, vnr_comm_t* c);
int vnr_renderer_detach_comm(vnr_renderer_t* r);

/* vnrMemoryQuery (api.h:186): bytes of device memory held by volumes / renderers */
int vnr_memory_query(size_t* used_by_renderer, size_t* used_by_network);

#ifdef __cplusplus
}
#endif
#endif /* VNR_C_H */
