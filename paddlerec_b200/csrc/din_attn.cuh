// DIN attention pooling kernels (K4)
#pragma once
#include "common.cuh"
