// planet_buffers.h -- the reference's geometry-buffer entry points, headless.
//
// render.h:80-81 declares `GLuint CreateVertexBuffer(unsigned int size, const void *data)` and
// `CreateIndexBuffer` (render.cpp:20-28: glGenBuffers + glBufferData), and InitPlanet hands them
// the patch mesh it built on the CPU (main.cpp:479-480).  Here a buffer object is a block of
// device memory: same names, same arguments, same "0 means failure" convention, and -- as with
// glBufferData -- `data` may be NULL when a kernel (K1, K3) is going to fill the buffer.
// MapBuffer() returns the CUDA pointer behind a handle, i.e. what
// cudaGraphicsResourceGetMappedPointer would return for a GL buffer under CUDA-GL interop.
// Header-only; include it in ONE translation unit of a host program that links cudart.
#ifndef PLANET_BUFFERS_H
#define PLANET_BUFFERS_H

#include <cuda_runtime_api.h>

#include <vector>

#ifndef PLANET_HOST_NO_GL_TYPES
typedef unsigned int GLuint;
#endif

struct DeviceBuffer { void *ptr; unsigned int size; };
static std::vector<DeviceBuffer> g_device_buffers;                   // handle = index + 1 (0 is GL's "no buffer")

static GLuint CreateDeviceBuffer(unsigned int size, const void *data)
{
    void *ptr = nullptr;
    if (cudaMalloc(&ptr, size ? size : 1) != cudaSuccess) return 0;
    if (data && cudaMemcpy(ptr, data, size, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(ptr); return 0; }
    g_device_buffers.push_back({ ptr, size });
    return (GLuint)g_device_buffers.size();
}

static inline GLuint CreateVertexBuffer(unsigned int size, const void *data) { return CreateDeviceBuffer(size, data); }   // render.h:80
static inline GLuint CreateIndexBuffer(unsigned int size, const void *data) { return CreateDeviceBuffer(size, data); }    // render.h:81

static inline void *MapBuffer(GLuint buffer)
{
    return (buffer == 0 || buffer > g_device_buffers.size()) ? nullptr : g_device_buffers[buffer - 1].ptr;
}

static inline void DeleteBuffers()                                   // glDeleteBuffers for everything created
{
    for (auto &b : g_device_buffers) cudaFree(b.ptr);
    g_device_buffers.clear();
}

#endif // PLANET_BUFFERS_H
