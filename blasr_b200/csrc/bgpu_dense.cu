#include "bgpu_common.cuh"
