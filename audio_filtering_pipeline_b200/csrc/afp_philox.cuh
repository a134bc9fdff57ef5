// Dither + quantise epilogue (kernel 5): Philox4x32-10 counter RNG, TPDF/RPDF dither, int16/int24.
//
// The reference's apply_dither (stream_process.py:61-64) adds +-1/2 LSB *float* noise from the
// unseeded numpy RNG and never quantises, so this stage is build-defined.  The
// arithmetic contract (bit for bit) is written down in oracle/dither_quant.py:
//   key     = (seed lo, seed hi)
//   counter = (c lo, c hi, stream_id, 0),  c = absolute_output_sample_index >> 2
//   sample i uses output word i & 3 of its block (four samples per Philox call - SURVEY.md 8(d));
//   a = word >> 16, b = word & 0xffff (two 16-bit uniforms)
//   TPDF d = float(a + b - (2^16-1)) * 2^-16        RPDF d = float(2*a + 1 - 2^16) * 2^-17
//   q = clip(rint(fl32(fl32(y*2^(bits-1)) + d)), -2^(bits-1), 2^(bits-1)-1)   (no fused multiply-add)
#pragma once
#include <cstdint>

namespace afp {

enum : int { FMT_F32 = 0, FMT_I16 = 1, FMT_I24_IN_I32 = 2, FMT_I24_PACKED = 3 };
enum : int { DITHER_NONE = 0, DITHER_TPDF = 1, DITHER_RPDF = 2 };

struct Philox4 { uint32_t w[4]; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    Philox4 o;
    o.w[0] = c0; o.w[1] = c1; o.w[2] = c2; o.w[3] = c3;
    return o;
}

// The same function for NC independent counters, round by round: the ten rounds of one call are a chain of dependent
// multiplies (ncu on the tensor kernel's epilogue warps: fixed-latency `wait` stalls, 0.2 instructions per clock and warp),
// NC chains side by side fill the gaps, and the round keys are computed once for all of them.  Bit-identical results.
template <int NC>
__device__ __forceinline__ void philox4x32_10_multi(uint32_t (&c0)[NC], uint32_t (&c1)[NC], uint32_t (&c2)[NC], uint32_t (&c3)[NC],
                                                    uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, c0[i]), lo0 = 0xD2511F53u * c0[i];
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2[i]), lo1 = 0xCD9E8D57u * c2[i];
            c0[i] = hi1 ^ c1[i] ^ k0;
            c1[i] = lo1;
            c2[i] = hi0 ^ c3[i] ^ k1;
            c3[i] = lo0;
        }
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

// Dither value in LSB units from one 32-bit word (two 16-bit uniforms).
// The integer (a+b or 2a+1, < 2^18) becomes a float by planting it in the mantissa of 2^23 - one integer
// add instead of an I2F conversion - and the offset and the power-of-two scale fold into one FMA whose
// result is exactly representable, so the value is the one the definition above states.
template <int MODE>
__device__ __forceinline__ float dither_lsb(uint32_t w) {
    const uint32_t a = w >> 16, b = w & 0xFFFFu;
    if (MODE == DITHER_TPDF) {
        const float t = __uint_as_float(0x4B000000u + a + b);                    // 2^23 + a + b
        return __fmaf_rn(t, 1.52587890625e-05f, -(8388608.0f + 65535.0f) * 1.52587890625e-05f);
    }
    const float t = __uint_as_float(0x4B000000u + 2u * a + 1u);                  // 2^23 + 2a + 1
    return __fmaf_rn(t, 7.62939453125e-06f, -(8388608.0f + 65536.0f) * 7.62939453125e-06f);
}

// float sample -> integer code of a BITS-bit signed format, d in LSB units.
template <int BITS>
__device__ __forceinline__ int quantise(float y, float d) {
    constexpr float scale = (float)(1 << (BITS - 1));
    constexpr int lim = 1 << (BITS - 1);
    const float v = __fadd_rn(__fmul_rn(y, scale), d);
    // round-half-even conversion (saturating; NaN -> 0), then the clip in the integer domain: the same
    // codes as clip(rint(v)) with three instructions fewer
    return min(max(__float2int_rn(v), -lim), lim - 1);
}

// Reference-style float dither (apply_dither(audio, bit_depth), stream_process.py:61-64): y + d * LSB with
// LSB = 2^-(bit_depth-1), both steps rounded to float32.
__device__ __forceinline__ float dither_float(float y, float d, float lsb) {
    return __fadd_rn(y, __fmul_rn(d, lsb));
}

}  // namespace afp
