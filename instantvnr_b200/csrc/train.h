// train.h -- launchers of the sampler / training / macrocell kernels
#pragma once
#include "volume.h"
namespace vnr {
void macrocell_update_explicit(Volume* v, const float* d_xyz, const float* d_values, size_t n, cudaStream_t s);
void macrocell_update_implicit(Volume* v, cudaStream_t s);
void macrocell_update_max_opacity(Volume* v, cudaStream_t s);
void sample_batch(Volume* v, float* d_xyz, float* d_target, size_t n, cudaStream_t s);
void sample_at(Volume* v, const float* d_xyz, float* d_out, size_t n, int hw_texture, cudaStream_t s);
void train_ensure_buffers(Volume* v);
void reset_optimizer_state(Volume* v);
void train_preload_kernels(const Volume* v);     // load every kernel of a training step before the first peer barrier is enqueued
void macrocell_preload_kernels();
void outofcore_preload_kernels();
int train_profile_words();
void train_grads(Volume* v, const float* d_xyz, const float* d_target, size_t n, size_t n_global, cudaStream_t s);
void optimizer_step(Volume* v, cudaStream_t s);
void dp_optimizer_step(Volume* v, cudaStream_t s);
void dp_finish_step(Volume* v, cudaStream_t s);
void load_groundtruth_file(Volume* v, const char* path, int type, uint64_t offset, bool big_endian, float vmin, float vmax, float* range_out);
void outofcore_open(Volume* v, const char* path, int type, uint64_t offset, float vmin, float vmax, uint32_t n_concurrent, uint32_t n_blocks);
void outofcore_release(Volume* v);
void outofcore_set_rank(Volume* v, int rank);   // data-parallel rank r selects (and re-reads) its own random slabs
void outofcore_sample(Volume* v, float* d_xyz, float* d_target, size_t n, cudaStream_t s);
void outofcore_info(Volume* v, uint32_t* n_slots, uint32_t* n_refresh, uint64_t* slot_bytes, uint64_t* first_voxel, uint32_t* length, uint64_t* bytes_uploaded);
constexpr int kSlicesPerBlob = 16;
void decode_progressive(Volume* v, cudaStream_t s);
cudaError_t launch_volume_samples(const float* vol, const int* dims3, const float4* samples, const float4* samples_alt, float* out,
                                  const uint32_t* n_dev, const uint32_t* round_dev, size_t n_max, cudaStream_t stream);
double volume_psnr(Volume* v, cudaStream_t s);
double volume_ssim(Volume* v, float* h_map, cudaStream_t s);
double volume_test_loss(Volume* v, size_t batch, cudaStream_t s);
void volume_export(Volume* v, const char* path, int which, float* range_out, cudaStream_t s);
void apply_l2_policy(const Volume* v, cudaStream_t s);
void train_side_stream(Volume* v);
void train_steps(Volume* v, int steps, size_t batch, bool update_macrocell, cudaStream_t s);
}
