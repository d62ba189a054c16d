// train.h -- launchers of the sampler / training / macrocell kernels
#pragma once
#include "volume.h"
namespace vnr {
}
