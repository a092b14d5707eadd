// stand-in for glad: only the GLuint type and the GL-interop registration symbol are referenced (VtBuffer.hpp L167-179)
#pragma once
#include <cuda_runtime.h>
typedef unsigned int GLuint;
struct cudaGraphicsResource;
#ifndef cudaGraphicsRegisterFlagsNone
#define cudaGraphicsRegisterFlagsNone 0
#endif
inline cudaError_t cudaGraphicsGLRegisterBuffer(struct cudaGraphicsResource**, GLuint, unsigned int) { return cudaErrorNotSupported; }
