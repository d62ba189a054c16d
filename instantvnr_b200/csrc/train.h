// train.h -- launchers of the sampler / training / macrocell kernels
#pragma once
#include "volume.h"
namespace vnr {
void macrocell_update_explicit(Volume* v, const float* d_xyz, const float* d_values, size_t n, cudaStream_t s);
void macrocell_update_implicit(Volume* v, cudaStream_t s);
void macrocell_update_max_opacity(Volume* v, cudaStream_t s);
}
