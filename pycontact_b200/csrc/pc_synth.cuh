// pc_synth.cuh -- on-device synthetic trajectory generator for bench.py.
// frame f = base + amp[res] * sin(2*pi*f/period[res] + phase[res]) + sigma * N(0,1),
// with a counter-based generator keyed (seed, frame, atom): any frame can be regenerated
// independently on any GPU, which is what lets ranks generate their own frame shards.
#pragma once
#include "pc_common.cuh"

__device__ __forceinline__ unsigned long long pc_splitmix64(unsigned long long x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256) pc_synth_kernel(const float *__restrict__ base, const int *__restrict__ res,
                                                       const float *__restrict__ amp, const float *__restrict__ period,
                                                       const float *__restrict__ phase, int natoms, float sigma,
                                                       unsigned long long seed, int frame0, int layout, long long pitch,
                                                       float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= natoms) return;
    const int f = frame0 + blockIdx.y;
    const int r = res[i];
    const float s = sinf(6.283185307179586f * (float)f / period[r] + phase[r]);
    const unsigned long long k0 = pc_splitmix64(seed ^ ((unsigned long long)(unsigned)f << 32) ^ (unsigned)i);
    const unsigned long long k1 = pc_splitmix64(k0);
    // four uniforms in (0,1] -> two Box-Muller pairs, three normals used
    const float u0 = ((unsigned)(k0 >> 40) + 1.0f) * (1.0f / 16777216.0f);
    const float u1 = ((unsigned)(k0 & 0xffffffu) + 1.0f) * (1.0f / 16777216.0f);
    const float u2 = ((unsigned)(k1 >> 40) + 1.0f) * (1.0f / 16777216.0f);
    const float u3 = ((unsigned)(k1 & 0xffffffu) + 1.0f) * (1.0f / 16777216.0f);
    const float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
    float sa, ca, sb, cb;
    sincosf(6.283185307179586f * u1, &sa, &ca);
    sincosf(6.283185307179586f * u3, &sb, &cb);
    (void)cb;
    const float nx = ra * ca, ny = ra * sa, nz = rb * sb;
    const float x = base[3 * i] + amp[3 * r] * s + sigma * nx;
    const float y = base[3 * i + 1] + amp[3 * r + 1] * s + sigma * ny;
    const float z = base[3 * i + 2] + amp[3 * r + 2] * s + sigma * nz;
    float *o = out + (long long)blockIdx.y * pitch;
    if (layout == 4) {
        reinterpret_cast<float4 *>(o)[i] = make_float4(x, y, z, 0.f);
    } else {
        o[3ll * i] = x; o[3ll * i + 1] = y; o[3ll * i + 2] = z;
    }
}

static inline void pc_synth_launch(const float *base, const int *res, const float *amp, const float *period,
                                   const float *phase, int natoms, float sigma, unsigned long long seed, int frame0,
                                   int nframes, int layout, long long pitch, float *out, cudaStream_t s) {
    // grid.y is limited to 65535 frames per launch
    for (int f = 0; f < nframes; f += 32768) {
        const int nf = nframes - f < 32768 ? nframes - f : 32768;
        dim3 grid((natoms + 255) / 256, nf);
        pc_synth_kernel<<<grid, 256, 0, s>>>(base, res, amp, period, phase, natoms, sigma, seed, frame0 + f, layout,
                                             pitch, out + (long long)f * pitch);
    }
}
