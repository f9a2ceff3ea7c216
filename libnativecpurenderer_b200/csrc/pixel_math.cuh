// Device-side scalar pieces shared by the kernels.  Every f64 operation goes through a round-to-nearest
// intrinsic so that no compiler flag can contract the reference's expression trees into FMAs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MUL(a, b) __dmul_rn((a), (b))
#define ADD(a, b) __dadd_rn((a), (b))
#define SUB(a, b) __dsub_rn((a), (b))
#define DIV(a, b) __ddiv_rn((a), (b))

// GetBufferAsUInt8, reference cpp:52-57: (iu8)(v * 255) as x86-64 gcc compiles it — cvttsd2si to a 32-bit
// integer (truncate toward zero; NaN / out-of-range give the "integer indefinite" 0x80000000), low byte kept.
__device__ __forceinline__ unsigned char ncr_to_u8(double v) {
    const double s = MUL(v, 255.0);
    int t;
    if (!(fabs(s) < 2147483648.0)) t = (int)0x80000000;
    else t = __double2int_rz(s);
    return (unsigned char)(t & 0xff);
}
