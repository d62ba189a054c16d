// render.h -- launchers of the marcher kernels (render.cu)
#pragma once
#include "volume.h"
namespace vnr {
}
