// Hand-written device functors for the jump-process models of the reference's tutorials / tests.
//
// Each functor restates one model's `compute_rates` + `apply_transitionings` pair for a single
// trajectory held in registers (int32 compartment counts):
//   propose / apply : one tau-leap step, split so that the sampling strategy lives in ONE place
//              (jump_run.cuh): `propose` lists the step's binomial draws (n_d, p_d) -- all of them are
//              functions of the state at the START of the step -- and `apply` books the drawn counts.
//              Bit d of DEP_MASK says draw d is conditional on draw d-1: n_d <- n_d - T_{d-1}.
//              The law is the reference's (`_tau_leap` + `_multinomial_rvs`,
//              /root/reference/src/pySODM/models/base.py:283-348): for a compartment with leave
//              probabilities p_0..p_{m-1} (p_i = 1 - exp(-tau*rate_i)) the counts are
//              Multinomial(n; [p_0..p_{m-1}, s] / W) with, when `quirk` (the reference's behaviour,
//              SURVEY F1), s = 1 - sum_{i<m-1} p_i  and  W = 1 + p_{m-1};  otherwise s = 1 - sum_i p_i,
//              W = 1.  It is drawn as sequential conditional binomials T_i ~ Bin(n - sum_{j<i} T_j,
//              p_i / (W - sum_{j<i} p_j)), which is the same multinomial whatever the peeling order.
//   propensities / fire : the channel list of `_SSA` (base.py:231-269) flattened in the reference's
//              order (state, transition, element) and the effect of one event of a channel.
// Rates that do not depend on the state are folded into per-trajectory constants by `prepare`.
#pragma once
#include "fdiv.cuh"
#include <cmath>
#include <cstdint>
#include "binomial.cuh"

namespace sodm {

enum : uint32_t {
    ST_OK = 0u,
    ST_BAD_PROB = 1u,        // a transition probability was NaN / outside [0,1] (NumPy would raise)
    ST_NONFINITE = 2u,       // non-finite model output
    ST_STEP_TOO_SMALL = 4u,  // ODE: step size underflow (scipy: "Required step size is less than spacing")
    ST_MAX_STEPS = 8u,       // step / event budget exhausted
    ST_TRUNCATED_Y0 = 16u,   // a non-integer initial count was truncated (base.py:345 truncates per step)
};

__host__ __device__ __forceinline__ float leave_prob(float rate, float tau) { return -expm1f(-tau * rate); }

// conditional probability of the single exit of a compartment
__host__ __device__ __forceinline__ float cond1(float p, bool quirk) { return quirk ? p / (1.0f + p) : p; }

// ------------------------------------------------------------------------------------------------
// SIR without dimensions; gamma is a rate.   /root/reference/src/tests/test_JumpProcess.py:11-27
// states S,I,R ; parameters beta,gamma ; channels S->I, I->R
// ------------------------------------------------------------------------------------------------
struct SirJump {
    static constexpr int NS = 3, NPAR = 2, NCH = 2, TDPF_OFFSET = 0, TDPF_COUNT = 0;
    struct Pre { float beta, cI; };

    __host__ __device__ static Pre prepare(const double* par, float tau, bool quirk) {
        Pre pre;
        pre.beta = (float)par[0];
        pre.cI = cond1(leave_prob((float)par[1], tau), quirk);
        return pre;
    }
    static constexpr int NDRAW = 2;
    static constexpr uint32_t DEP_MASK = 0u;
    template <class W>
    __host__ __device__ static void propose(const int* y, const Pre& pre, const float*, float tau, bool quirk, W& w) {
        const float tot = (float)y[0] + (float)y[1] + (float)y[2];
        const float pS = leave_prob(pre.beta * ((float)y[1] / tot), tau);
        w.put(0, y[0], cond1(pS, quirk));
        w.put(1, y[1], pre.cI);
    }
    template <class W>
    __host__ __device__ static void apply(int* y, const W& w) {
        const int si = w.get(0), ir = w.get(1);
        y[0] -= si;
        y[1] += si - ir;
        y[2] += ir;
    }
    __host__ __device__ static void propensities(const int* y, const double* par, double* a) {
        const double S = y[0], I = y[1], R = y[2];
        a[0] = S * (par[0] * (I / (S + I + R)));
        a[1] = I * par[1];
    }
    __host__ __device__ static void fire(int* y, int ch) {
        if (ch == 0) { y[0] -= 1; y[1] += 1; }
        else         { y[1] -= 1; y[2] += 1; }
    }
};

// ------------------------------------------------------------------------------------------------
// Vector-borne SIR (4 age groups) + SI (vector).   /root/reference/tutorials/SIR_SI/models.py:4-40
// flat states: S[0:4] I[4:8] R[8:12] S_v[12] I_v[13]
// flat params: alpha[0:4] beta[4] gamma[5]
// channels   : S->I (4), I->R (4), S_v->I_v (1), S_v death (1, sampled, never applied), I_v->S_v (1)
// ------------------------------------------------------------------------------------------------
struct SirSiJump {
    static constexpr int NS = 14, NPAR = 6, NCH = 11, TDPF_OFFSET = 0, TDPF_COUNT = 0;
    struct Pre { float alpha[4]; float cI, pDeath, cIv; };

    __host__ __device__ static Pre prepare(const double* par, float tau, bool quirk) {
        Pre pre;
#pragma unroll
        for (int i = 0; i < 4; ++i) pre.alpha[i] = (float)par[i];
        pre.cI = cond1(leave_prob(1.0f / (float)par[5], tau), quirk);
        pre.pDeath = leave_prob(1.0f / (float)par[4], tau);
        pre.cIv = cond1(pre.pDeath, quirk);
        return pre;
    }
    // draws: 0-3 S_a->I_a, 4-7 I_a->R_a, 8 S_v->I_v, 9 I_v->S_v.  S_v's second exit (natural death) is
    // sampled by the reference but never applied (models.py:37-38): it is not drawn here, yet -- being S_v's
    // LAST exit -- its probability sets the normaliser W = 1 + p_death of the first one (quirk).
    static constexpr int NDRAW = 10;
    static constexpr uint32_t DEP_MASK = 0u;
    template <class W>
    __host__ __device__ static void propose(const int* y, const Pre& pre, const float*, float tau, bool quirk, W& w) {
        const float frac_v = (float)y[13] / ((float)y[12] + (float)y[13]);
        float force = 0.0f;                       // sum_a alpha_a * I_a / N_a  (rate of S_v -> I_v)
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float Na = (float)y[a] + (float)y[4 + a] + (float)y[8 + a];
            force += pre.alpha[a] * ((float)y[4 + a] / Na);
            w.put(a, y[a], cond1(leave_prob(pre.alpha[a] * frac_v, tau), quirk));
            w.put(4 + a, y[4 + a], pre.cI);
        }
        const float p0 = leave_prob(force, tau);
        const float Wn = quirk ? 1.0f + pre.pDeath : 1.0f;
        w.put(8, y[12], p0 / Wn);
        w.put(9, y[13], pre.cIv);
    }
    template <class W>
    __host__ __device__ static void apply(int* y, const W& w) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int si = w.get(a), ir = w.get(4 + a);
            y[a] -= si;
            y[4 + a] += si - ir;
            y[8 + a] += ir;
        }
        const int sv_iv = w.get(8), iv_sv = w.get(9);
        y[12] += iv_sv - sv_iv;
        y[13] += sv_iv - iv_sv;
    }
    __host__ __device__ static void propensities(const int* y, const double* par, double* a) {
        const double Sv = y[12], Iv = y[13];
        const double frac_v = div_z(Iv, Sv + Iv);                     // zero counts stay on the division's fast path (fdiv.cuh)
        double force = 0.0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const double S = y[g], I = y[4 + g], R = y[8 + g];
            a[g] = S * (par[g] * frac_v);
            a[4 + g] = I * (1.0 / par[5]);
            force += par[g] * div_z(I, S + I + R);
        }
        a[8] = Sv * force;
        a[9] = Sv * (1.0 / par[4]);
        a[10] = Iv * (1.0 / par[4]);
    }
    __host__ __device__ static void fire(int* y, int ch) {
        // static indices only, so that y[] stays in registers
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            if (ch == g)     { y[g] -= 1;     y[4 + g] += 1; }     // S_a -> I_a
            if (ch == 4 + g) { y[4 + g] -= 1; y[8 + g] += 1; }     // I_a -> R_a
        }
        if (ch == 8)  { y[12] -= 1; y[13] += 1; }
        /* ch == 9: S_v death -- drawn, never applied (models.py:37-38) */
        if (ch == 10) { y[13] -= 1; y[12] += 1; }
    }
};

// ------------------------------------------------------------------------------------------------
// Influenza 2017-18: age-structured SEIR with undetected carriers.
// /root/reference/tutorials/influenza_1718/models.py:32-69
// flat states: S[0:4] E[4:8] Ip[8:12] Iud[12:16] Id[16:20] R[20:24] Im_inc[24:28]
// flat params: alpha[0] beta[1] gamma[2] delta[3] f_ud[4:8] N[8:24] (row-major 4x4, time dependent)
// channels   : S->E (4), E->Ip (4), Ip->Iud (4), Ip->Id (4), Iud->R (4), Id->R (4)
// ------------------------------------------------------------------------------------------------
// Slot of draw `kind` (0 S->E, 1 E->Ip, 2 Ip->Iud, 3 Ip->Id, 4 Iud->R, 5 Id->R) of age group a: declaration order.  The
// order in which a lane WALKS its draws is not the functor's business: every lane sorts its draws by its own first-step
// means (reorder_draws, jump_run.cuh) -- the generic rule that serves generated models too.  (Round 1 shipped a table
// ordered by this model's season-average means; it measured the same as the generic sort, 125.8 vs 125.6 ms, and was
// fitted to the benchmark's own trajectory, so it is gone.)
__host__ __device__ constexpr int flu_slot(int a, int kind) { return 6 * a + kind; }

struct InfluenzaJump {
    static constexpr int NS = 28, NPAR = 24, NCH = 24, TDPF_OFFSET = 8, TDPF_COUNT = 16;
    struct Pre { float beta, cE, cD; float cIp0[4], cIp1[4]; };

    __host__ __device__ static Pre prepare(const double* par, float tau, bool quirk) {
        Pre pre;
        pre.beta = (float)par[1];
        pre.cE = cond1(leave_prob(1.0f / (float)par[0], tau), quirk);
        pre.cD = cond1(leave_prob(1.0f / (float)par[3], tau), quirk);
        const float inv_gamma = 1.0f / (float)par[2];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float f = (float)par[4 + a];
            const float p0 = leave_prob(f * inv_gamma, tau);            // Ip -> Iud
            const float p1 = leave_prob((1.0f - f) * inv_gamma, tau);   // Ip -> Id (last exit)
            const float W = quirk ? 1.0f + p1 : 1.0f;
            pre.cIp0[a] = p0 / W;
            pre.cIp1[a] = p1 / (W - p0);
        }
        return pre;
    }
    // 24 draws: per age group a  S->E, E->Ip, Ip->Iud, Ip->Id (conditional on Ip->Iud), Iud->R, Id->R  (slots: flu_slot)
    static constexpr int NDRAW = 24;
    __host__ __device__ static constexpr int slot(int a, int kind) { return flu_slot(a, kind); }
    static constexpr uint32_t DEP_MASK = (1u << flu_slot(0, 3)) | (1u << flu_slot(1, 3)) | (1u << flu_slot(2, 3)) | (1u << flu_slot(3, 3));
    // `Nmat` = the contact matrix in force during this step (16 floats, row-major).
    template <class W>
    __host__ __device__ static void propose(const int* y, const Pre& pre, const float* Nmat, float tau, bool quirk, W& w) {
        float x[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float T = (float)(y[a] + y[4 + a] + y[8 + a] + y[12 + a] + y[16 + a] + y[20 + a]);
            x[a] = ((float)y[8 + a] + (float)y[12 + a] + 0.22f * (float)y[16 + a]) / T;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float foi = pre.beta * (Nmat[4 * a] * x[0] + Nmat[4 * a + 1] * x[1] +
                                          Nmat[4 * a + 2] * x[2] + Nmat[4 * a + 3] * x[3]);
            w.put(slot(a, 0), y[a], cond1(leave_prob(foi, tau), quirk));
            w.put(slot(a, 1), y[4 + a], pre.cE);
            w.put(slot(a, 2), y[8 + a], pre.cIp0[a]);
            w.put(slot(a, 3), y[8 + a], pre.cIp1[a]);
            w.put(slot(a, 4), y[12 + a], pre.cD);
            w.put(slot(a, 5), y[16 + a], pre.cD);
        }
    }
    template <class W>
    __host__ __device__ static void apply(int* y, const W& w) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int s_e = w.get(slot(a, 0)), e_ip = w.get(slot(a, 1)), ip_iud = w.get(slot(a, 2));
            const int ip_id = w.get(slot(a, 3)), iud_r = w.get(slot(a, 4)), id_r = w.get(slot(a, 5));
            y[a] -= s_e;
            y[4 + a] += s_e - e_ip;
            y[8 + a] += e_ip - ip_iud - ip_id;
            y[12 + a] += ip_iud - iud_r;
            y[16 + a] += ip_id - id_r;
            y[20 + a] += iud_r + id_r;
            y[24 + a] = ip_id;
        }
    }
    // SSA (base.py:231-269): channels in the reference's order (state, transition, element) = S->E[4], E->Ip[4], Ip->Iud[4],
    // Ip->Id[4], Iud->R[4], Id->R[4]; propensity = state * rate with the rates of models.py:47-55 in FP64.  par[8:24] holds
    // the contact matrix in force at the event time (ssa_trajectory copies the table row there).
    __host__ __device__ static void propensities(const int* y, const double* par, double* a) {
        double x[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const double T = (double)y[g] + (double)y[4 + g] + (double)y[8 + g] + (double)y[12 + g] + (double)y[16 + g] + (double)y[20 + g];
            x[g] = div_z((double)y[8 + g] + (double)y[12 + g] + 0.22 * (double)y[16 + g], T);
        }
        const double e_rate = 1.0 / par[0], inv_gamma = 1.0 / par[2], d_rate = 1.0 / par[3];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const double* N = par + 8 + 4 * g;
            const double foi = par[1] * (N[0] * x[0] + N[1] * x[1] + N[2] * x[2] + N[3] * x[3]);
            a[g] = (double)y[g] * foi;
            a[4 + g] = (double)y[4 + g] * e_rate;
            a[8 + g] = (double)y[8 + g] * (par[4 + g] * inv_gamma);
            a[12 + g] = (double)y[8 + g] * ((1.0 - par[4 + g]) * inv_gamma);
            a[16 + g] = (double)y[12 + g] * d_rate;
            a[20 + g] = (double)y[16 + g] * d_rate;
        }
    }
    // one event of channel ch (apply_transitionings with a one-hot transitionings dict, models.py:59-69); Im_inc is the
    // Ip->Id count of THIS event: 1 for the age group that fired it, 0 elsewhere.  Static indices keep y[] in registers.
    __host__ __device__ static void fire(int* y, int ch) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            if (ch == g)      { y[g] -= 1;      y[4 + g] += 1; }
            if (ch == 4 + g)  { y[4 + g] -= 1;  y[8 + g] += 1; }
            if (ch == 8 + g)  { y[8 + g] -= 1;  y[12 + g] += 1; }
            if (ch == 12 + g) { y[8 + g] -= 1;  y[16 + g] += 1; }
            if (ch == 16 + g) { y[12 + g] -= 1; y[20 + g] += 1; }
            if (ch == 20 + g) { y[16 + g] -= 1; y[20 + g] += 1; }
            y[24 + g] = ch == 12 + g ? 1 : 0;
        }
    }
};

}  // namespace sodm
