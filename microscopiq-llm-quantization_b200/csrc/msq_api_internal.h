// msq_api_internal.h -- includes shared by the translation units of libmsq_b200.so
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string.h>
