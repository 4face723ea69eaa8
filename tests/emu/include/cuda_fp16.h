// TEST INFRASTRUCTURE: stands in for the CUDA toolkit header of the same name when the product sources are compiled for the CPU (tests/emu)
#pragma once
#include "../cuda_on_host.h"
