// FP64 division whose ZERO numerators stay on the fast path.
//
// CUDA's IEEE double division is an inline sequence (MUFU.RCP64H + Newton steps + a range check on the quotient) with an
// out-of-line slow path for quotients that are zero, subnormal-range, infinite or NaN.  A 0 / x is such a quotient: every
// lane whose numerator is exactly zero calls the slow path (~60 instructions, executed with only those lanes active).  The
// models on this path divide zeros all the time -- the error estimate of a component that does not move, a reaction rate
// where there is no substrate yet, an infected fraction of an age group without infected -- and ncu showed the consequence:
// 35 % of the reactor kernel's executed instructions were that slow path at 7 of 32 lanes (profiles/r2_pfr25_ncu_lines.txt;
// the lanes of a warp that hold no reactor node contributed 0 / 0 on top -- they now carry a harmless non-zero state,
// sodm_ode.cu).  `div_z` divides 1 / b in those lanes instead and multiplies the (signed) zero back in: the result is
// bit-identical to a / b -- +-0 with the sign of a xor b -- for every finite or infinite b != 0; 0 / 0 and 0 / NaN still take
// the real division (NaN, as IEEE and the reference have it).  Scaling TINY numerators as well (below 2^-800, by 2^400 and
// back) was measured and dropped: 215 against 208 ms on the reactor -- the dust ahead of the concentration front is rare.
#pragma once

namespace sodm {

__host__ __device__ __forceinline__ double div_z(double a, double b) {
    const bool zero = (a == 0.0) & (b != 0.0) & (b == b);
    const double q = (zero ? 1.0 : a) / b;
    return zero ? a * q : q;
}

// The interpolation onto a grid time is needed on the few steps / events that cross one, but it is pure arithmetic on values
// that do not change inside the crossing loop, so the compiler hoists it in front of the loop and evaluates it -- one FP64
// division per state -- at EVERY step or event (ncu on the SSA kernel: 14 divisions per event, all through the slow path, 40 %
// of the kernel's instructions).  Passing the step's start time through an opaque asm statement inside the loop keeps the
// arithmetic where it belongs.
__host__ __device__ __forceinline__ double pin_in_loop(double x) {
#ifdef __CUDA_ARCH__
    asm volatile("" : "+d"(x));
#endif
    return x;
}

}  // namespace sodm
