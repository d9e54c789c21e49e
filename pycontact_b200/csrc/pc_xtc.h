// pc_xtc.h -- GROMACS XTC trajectory frames -> float32 coordinates in Angstrom (host code, no CUDA).
//
// Ingest step in front of the path: the reference reads trajectories through MDAnalysis
// (PyContact/core/ContactAnalyzer.py:282, core/multi_trajectory.py:362; its own tests open
// exampleData/md_noPBC.xtc, tests/test_basic.py:28-30), whose XTC reader is the xdrfile library.  xdrfile is not part of
// the reference tree; this is a restatement of its published frame layout and of its lossy integer coordinate
// compression (xdr3dfcoord): big-endian XDR words
//     magic 1995 | natoms | step | time | box[9] | natoms | precision | minint[3] | maxint[3] | smallidx | nbytes | bits...
// with <= 9 atoms stored as plain floats.  Every atom is three integers (coordinate * precision, rounded); the first
// atom of a run is coded in as many bits as the frame's integer bounding box needs (the three values packed as ONE
// mixed-radix number), followed by a flag bit and, if set, five bits giving the run length and whether the "small"
// range grows or shrinks; atoms inside a run are coded as differences to their predecessor in `smallidx` bits (three
// values packed with radix magicints[smallidx]), and the first two atoms of a run are swapped (water: O H H).
// Positions come back in nm; like MDAnalysis they are scaled by 10 to Angstrom in float32.
#pragma once
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

struct pc_xtc {
    FILE *fh = nullptr;
    int natoms = 0;
    std::vector<long long> offset;      // byte offset of every frame
    std::vector<unsigned char> raw;     // one frame's bytes
    std::vector<float> xyz;             // one frame, all atoms, nm
};

namespace pcxtc {
static const int magicints[] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256, 322, 406, 512,
                                645, 812, 1024, 1290, 1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003, 16384, 20642, 26007,
                                32768, 41285, 52015, 65536, 82570, 104031, 131072, 165140, 208063, 262144, 330280, 416127, 524287,
                                660561, 832255, 1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304, 5284491, 6658042,
                                8388607, 10568983, 13316085, 16777216};
constexpr int FIRSTIDX = 9;
constexpr int LASTIDX = (int)(sizeof(magicints) / sizeof(magicints[0]));

inline uint32_t be32(const unsigned char *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline float bef(const unsigned char *p) { const uint32_t u = be32(p); float f; memcpy(&f, &u, 4); return f; }

struct BitReader {
    const unsigned char *p;
    size_t n, cnt = 0;
    unsigned lastbits = 0, lastbyte = 0;
    bool overrun = false;
    unsigned next() { if (cnt >= n) { overrun = true; return 0; } return p[cnt++]; }
    unsigned bits(int nbits) {
        const unsigned mask = nbits >= 32 ? 0xffffffffu : (1u << nbits) - 1u;
        unsigned num = 0;
        while (nbits >= 8) {
            lastbyte = (lastbyte << 8) | next();
            num |= (lastbyte >> lastbits) << (nbits - 8);
            nbits -= 8;
        }
        if (nbits > 0) {
            if (lastbits < (unsigned)nbits) { lastbits += 8; lastbyte = (lastbyte << 8) | next(); }
            lastbits -= nbits;
            num |= (lastbyte >> lastbits) & ((1u << nbits) - 1u);
        }
        return num & mask;
    }
    // three integers packed as one mixed-radix number of `nbits` bits (radices sizes[0..2])
    void ints(int nbits, const unsigned sizes[3], int nums[3]) {
        unsigned bytes[32];
        int nbytes = 0;
        bytes[1] = bytes[2] = bytes[3] = 0;
        while (nbits > 8) { bytes[nbytes++] = bits(8); nbits -= 8; }
        if (nbits > 0) bytes[nbytes++] = bits(nbits);
        for (int i = 2; i > 0; i--) {
            unsigned num = 0;
            for (int j = nbytes - 1; j >= 0; j--) {
                num = (num << 8) | bytes[j];
                const unsigned q = num / sizes[i];
                bytes[j] = q;
                num -= q * sizes[i];
            }
            nums[i] = (int)num;
        }
        nums[0] = (int)(bytes[0] | (bytes[1] << 8) | (bytes[2] << 16) | (bytes[3] << 24));
    }
};

inline int sizeofint(unsigned size) {
    unsigned num = 1;
    int nbits = 0;
    while (size >= num && nbits < 32) { nbits++; num <<= 1; }
    return nbits;
}
inline int sizeofints(const unsigned sizes[3]) {
    unsigned bytes[32], nbytes = 1;
    bytes[0] = 1;
    for (int i = 0; i < 3; i++) {
        unsigned tmp = 0, k = 0;
        for (; k < nbytes; k++) { tmp = bytes[k] * sizes[i] + tmp; bytes[k] = tmp & 0xff; tmp >>= 8; }
        while (tmp != 0) { bytes[k++] = tmp & 0xff; tmp >>= 8; }
        nbytes = k;
    }
    unsigned num = 1;
    int nbits = 0;
    nbytes--;
    while (bytes[nbytes] >= num) { nbits++; num *= 2; }
    return nbits + (int)nbytes * 8;
}

// bytes of the frame that starts with `hdr` (92 header bytes read; fewer are looked at for <= 9 atoms); 0 = not a frame
inline long long frame_bytes(const unsigned char *hdr, int &natoms) {
    if (be32(hdr) != 1995u) return 0;
    natoms = (int)be32(hdr + 4);
    if (natoms <= 0 || (int)be32(hdr + 52) != natoms) return 0;
    if (natoms <= 9) return 56 + 12ll * natoms;
    const long long nbytes = (long long)be32(hdr + 88);
    return 92 + ((nbytes + 3) & ~3ll);
}

// one frame (raw bytes) -> xyz (nm), step, time, box (nm); returns an empty string or what is wrong
inline std::string decode(const unsigned char *raw, size_t nraw, int natoms, float *xyz, int *step, float *time, float *box) {
    if (step) *step = (int)be32(raw + 8);
    if (time) *time = bef(raw + 12);
    if (box) for (int k = 0; k < 9; k++) box[k] = bef(raw + 16 + 4 * k);
    if (natoms <= 9) {
        for (int k = 0; k < 3 * natoms; k++) xyz[k] = bef(raw + 56 + 4 * k);
        return "";
    }
    const float precision = bef(raw + 56);
    if (!(precision > 0.f)) return "bad precision";
    int minint[3], maxint[3];
    for (int k = 0; k < 3; k++) { minint[k] = (int)be32(raw + 60 + 4 * k); maxint[k] = (int)be32(raw + 72 + 4 * k); }
    int smallidx = (int)be32(raw + 84);
    const size_t nbytes = be32(raw + 88);
    if (92 + nbytes > nraw) return "truncated frame";
    if (smallidx < FIRSTIDX || smallidx >= LASTIDX) return "bad smallidx";
    unsigned sizeint[3], bitsizeint[3] = {0, 0, 0};
    for (int k = 0; k < 3; k++) {
        if (maxint[k] < minint[k]) return "bad integer bounds";
        sizeint[k] = (unsigned)(maxint[k] - minint[k]) + 1u;
    }
    int bitsize;
    if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffffu) {
        for (int k = 0; k < 3; k++) bitsizeint[k] = (unsigned)sizeofint(sizeint[k]);
        bitsize = 0;
    } else {
        bitsize = sizeofints(sizeint);
    }
    int smaller = magicints[smallidx - 1 > FIRSTIDX ? smallidx - 1 : FIRSTIDX] / 2;
    int smallnum = magicints[smallidx] / 2;
    unsigned sizesmall[3] = {(unsigned)magicints[smallidx], (unsigned)magicints[smallidx], (unsigned)magicints[smallidx]};
    BitReader br{raw + 92, nbytes};
    const float inv = 1.0f / precision;
    int i = 0, run = 0;
    float *out = xyz;
    while (i < natoms) {
        int cur[3], prev[3];
        if (bitsize == 0) {
            for (int k = 0; k < 3; k++) cur[k] = (int)br.bits((int)bitsizeint[k]);
        } else {
            br.ints(bitsize, sizeint, cur);
        }
        i++;
        for (int k = 0; k < 3; k++) { cur[k] += minint[k]; prev[k] = cur[k]; }
        const unsigned flag = br.bits(1);
        int is_smaller = 0;
        if (flag == 1) {
            run = (int)br.bits(5);
            is_smaller = run % 3;
            run -= is_smaller;
            is_smaller--;
        }
        if (run > 0) {
            if (i + run / 3 > natoms) return "run past the last atom";
            for (int k = 0; k < run; k += 3) {
                br.ints(smallidx, sizesmall, cur);
                i++;
                for (int c = 0; c < 3; c++) cur[c] += prev[c] - smallnum;
                if (k == 0) {
                    // the first two atoms of a run are stored swapped (water: hydrogen before oxygen compresses better)
                    for (int c = 0; c < 3; c++) { const int t = cur[c]; cur[c] = prev[c]; prev[c] = t; }
                    for (int c = 0; c < 3; c++) *out++ = (float)prev[c] * inv;
                } else {
                    for (int c = 0; c < 3; c++) prev[c] = cur[c];
                }
                for (int c = 0; c < 3; c++) *out++ = (float)cur[c] * inv;
            }
        } else {
            for (int c = 0; c < 3; c++) *out++ = (float)cur[c] * inv;
        }
        smallidx += is_smaller;
        if (smallidx < FIRSTIDX || smallidx >= LASTIDX) return "smallidx left its table";
        if (is_smaller < 0) {
            smallnum = smaller;
            smaller = smallidx > FIRSTIDX ? magicints[smallidx - 1] / 2 : 0;
        } else if (is_smaller > 0) {
            smaller = smallnum;
            smallnum = magicints[smallidx] / 2;
        }
        sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)magicints[smallidx];
        if (br.overrun) return "bit stream shorter than the frame's atoms";
    }
    // a correct parse ends inside the last byte of the stream (padding to a 4-byte word follows)
    if (br.cnt + 1 < nbytes) return "bit stream longer than the frame's atoms";
    return "";
}
}   // namespace pcxtc
