// pc_stream.cuh -- streaming packed float32 xyz coordinates from HBM through shared memory with the
// bulk-copy engine (cp.async.bulk, "1-D TMA") and mbarriers.
//
// The path's input is the (frames, atoms, 3) float32 array that sel.positions yields in the reference
// (core/multi_trajectory.py:401-417): 12 bytes per atom, frames 4-byte aligned only.  Every atom has
// to be looked at once per frame, so this read IS the HBM roofline of the large configurations.  A
// CTA walks a byte segment in CHUNK-byte pieces through a STAGES-deep ring: one thread issues the
// bulk copies (no register staging, no L1 pollution, up to STAGES*CHUNK bytes in flight per CTA),
// everybody else only ever reads shared memory, with stride-3-word accesses that are bank-conflict
// free.
//
// Alignment: bulk copies need 16-byte aligned addresses and sizes.  Chunk q of a segment [sb, se)
// copies the aligned window [g0, g0 + CHUNK + 16) with g0 = (sb & ~15) + q*CHUNK, clipped to the
// 16-byte granule that holds se - 1.  It may read up to 12 bytes before sb and after se, but never
// leaves a 16-byte granule that holds valid data, so it cannot fault.  Atoms are owned by the chunk
// that holds their first byte; the 16 bytes of overlap keep an owned atom's 12 bytes inside the slot.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t pc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void pc_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void pc_mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void pc_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pc_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool pc_mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(pc_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Spins until the phase with the given parity has completed.  Bounded: a copy that never lands (a
// bug, not a data condition) returns false after ~1 s instead of hanging the GPU.
__device__ __forceinline__ bool pc_mbar_wait(uint64_t *bar, uint32_t parity) {
    for (uint32_t spin = 0; spin < (1u << 24); spin++)
        if (pc_mbar_try_wait(bar, parity)) return true;
    return false;
}
__device__ __forceinline__ void pc_bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     pc_smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(pc_smem_u32(bar))
                 : "memory");
}

// L2 eviction-priority hints for the bulk copies (createpolicy + .L2::cache_hint).  The 1 M-atom configuration streams
// 11.4 MB of selection B per frame exactly once (evict_first: do not keep it), while selection A's 0.6 MB per frame is
// read twice -- bounding boxes, then the per-frame kernel -- with B's stream in between (evict_last: keep it).
__device__ __forceinline__ uint64_t pc_l2_policy(int kind) {          // 1 = evict_first, 2 = evict_last
    uint64_t p;
    if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void pc_bulk_g2s_hint(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     pc_smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(pc_smem_u32(bar)), "l"(policy)
                 : "memory");
}

// One contiguous run of atoms (packed xyz) split into chunks.
struct PcSegment {
    const unsigned char *sb;   // first byte of atom 0 of the segment
    int natoms;
    int atom0;                 // index (in its selection) of the segment's first atom
};

template <int CHUNK>
__device__ __forceinline__ int pc_seg_chunks(const PcSegment &s) {
    if (s.natoms <= 0) return 0;
    const unsigned long long lead = (unsigned long long)(uintptr_t)s.sb & 15ull;
    return (int)((lead + 12ull * (unsigned long long)s.natoms + CHUNK - 1) / CHUNK);
}

// Geometry of chunk q of a segment: what to copy and which atoms it owns.
struct PcChunk {
    const unsigned char *g0;   // 16-byte aligned source
    uint32_t bytes;            // multiple of 16, <= CHUNK + 16
    int k0, k1;                // owned atoms [k0, k1) relative to the segment
    int off0;                  // byte offset of atom k0 inside the slot
};

template <int CHUNK>
__device__ __forceinline__ PcChunk pc_chunk_of(const PcSegment &s, int q) {
    PcChunk c;
    const uintptr_t sb = (uintptr_t)s.sb, se = sb + 12ull * (unsigned long long)s.natoms;
    const uintptr_t g0 = (sb & ~(uintptr_t)15) + (uintptr_t)q * CHUNK;
    const uintptr_t gend = (se + 15) & ~(uintptr_t)15;
    uintptr_t g1 = g0 + CHUNK + 16;
    if (g1 > gend) g1 = gend;
    c.g0 = (const unsigned char *)g0;
    c.bytes = (uint32_t)(g1 - g0);
    // atoms whose first byte lies in [max(g0, sb), g0 + CHUNK)
    c.k0 = q == 0 ? 0 : (int)((g0 - sb + 11) / 12);
    const unsigned long long lim = (unsigned long long)(g0 + CHUNK - sb);
    c.k1 = (int)((lim + 11) / 12);
    if (c.k1 > s.natoms) c.k1 = s.natoms;
    c.off0 = (int)(sb + 12ull * (unsigned long long)c.k0 - g0);
    return c;
}

// Ring of STAGES slots of CHUNK + 16 bytes, one mbarrier each.  `use` counts chunks consumed by this
// CTA since the ring was initialised (slot = use % STAGES, parity = (use / STAGES) & 1).
template <int STAGES, int CHUNK>
struct PcRing {
    static constexpr int SLOT = CHUNK + 16;
    unsigned char *buf;        // STAGES * SLOT bytes, 16-byte aligned
    uint64_t *bar;             // STAGES barriers
    __device__ __forceinline__ unsigned char *slot(unsigned use) const { return buf + (size_t)(use % STAGES) * SLOT; }
    __device__ __forceinline__ void issue(unsigned use, const PcChunk &c) const {
        uint64_t *b = bar + (use % STAGES);
        pc_mbar_expect_tx(b, c.bytes);
        pc_bulk_g2s(slot(use), c.g0, c.bytes, b);
    }
    __device__ __forceinline__ bool wait(unsigned use) const { return pc_mbar_wait(bar + (use % STAGES), (use / STAGES) & 1u); }
};
