// Hand-written device functors for the ODE models of the reference's tutorials / tests: each restates
// one model's `integrate(t, *states, *parameters)` for a single trajectory, FP64 throughout (the parity
// bar is rtol 1e-6 against scipy's FP64 RK45).
#pragma once
#include "fdiv.cuh"
#include <cmath>
#include <cstdint>

namespace sodm {

// Tutorial SIR; gamma is a DURATION.   /root/reference/tutorials/SIR/workflow_tutorial.py:57-67
struct SirOde {
    static constexpr int N = 3, NPAR = 2;
    __host__ __device__ static void rhs(double, const double* y, const double* par, double* dy) {
        const double S = y[0], I = y[1], R = y[2];
        const double beta = par[0], gamma = par[1];
        const double Ntot = S + I + R;
        const double inf = beta * S * I / Ntot;
        const double rec = 1.0 / gamma * I;
        dy[0] = -inf;
        dy[1] = inf - rec;
        dy[2] = rec;
    }
};

// Test-suite SIR; gamma is a RATE.   /root/reference/src/tests/test_ODE.py:18-25
struct SirRateOde {
    static constexpr int N = 3, NPAR = 2;
    __host__ __device__ static void rhs(double, const double* y, const double* par, double* dy) {
        const double S = y[0], I = y[1], R = y[2];
        const double beta = par[0], gamma = par[1];
        const double Ntot = S + I + R;
        const double inf = beta * I * S / Ntot;
        dy[0] = -inf;
        dy[1] = inf - gamma * I;
        dy[2] = gamma * I;
    }
};

// Vector-borne SIR-SI.   /root/reference/tutorials/SIR_SI/models.py:53-66
// flat states S[0:4] I[4:8] R[8:12] S_v[12] I_v[13]; params alpha[0:4] beta[4] gamma[5]
struct SirSiOde {
    static constexpr int N = 14, NPAR = 6;
    __host__ __device__ static void rhs(double, const double* y, const double* par, double* dy) {
        const double S_v = y[12], I_v = y[13];
        const double beta = par[4], gamma = par[5];
        const double N_v = S_v + I_v;
        const double frac_v = I_v / N_v;
        double force = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const double S = y[a], I = y[4 + a], R = y[8 + a];
            const double Ntot = S + I + R;
            const double inf = par[a] * frac_v * S;
            dy[a] = -inf;
            dy[4 + a] = inf - 1.0 / gamma * I;
            dy[8 + a] = 1.0 / gamma * I;
            force += par[a] * (I / Ntot) * S_v;
        }
        dy[12] = -force + (1.0 / beta) * N_v - (1.0 / beta) * S_v;
        dy[13] = force - (1.0 / beta) * I_v;
    }
};

// Batch enzyme kinetics S + A <-> Es + W.   /root/reference/tutorials/enzyme_kinetics/models.py:62-69
// states S,A,Es,W ; params c_enzyme, Vf_Ks, R_AS, R_AW, R_Es, K_eq
struct PpbbOde {
    static constexpr int N = 4, NPAR = 6;
    __host__ __device__ static void rhs(double, const double* y, const double* par, double* dy) {
        const double S = y[0], A = y[1], Es = y[2], W = y[3];
        const double v = par[0] * div_z(par[1] * (S * A - (1.0 / par[5]) * Es * W), A + par[2] * S + par[3] * W + par[4] * Es);
        dy[0] = -v;
        dy[1] = -v;
        dy[2] = v;
        dy[3] = v;
    }
};

// 1-D packed-bed plug-flow reactor, one axial node.   /root/reference/tutorials/enzyme_kinetics/models.py:17-51
// params: delta_x[0] epsilon[1] u[2] rho_B[3] Vf_Ks[4] R_AS[5] R_AW[6] K_eq[7] R_Es[8] kL_a[9:13] D_ax[13:17]
// Node 0 is the Dirichlet inlet (all derivatives 0, the loop in the reference starts at j = 1); the outlet node
// mirrors its left neighbour in the dispersion stencil; the reaction runs on the solid phase only.
// cl / cm / cr: fluid concentrations of the 4 species at nodes j-1, j, j+1 (cr ignored at the outlet);
// cs: solid concentrations at node j.
struct PfrNode {
    static constexpr int NPAR = 17;
    // Per-trajectory coefficients, formed ONCE with the reference's own quotients (D_ax/dx^2, u/dx, kL_a/eps,
    // kL_a/(1-eps), 1/K_eq -- `models.py:37-49`), so that a right-hand-side evaluation has one FP64 division (the
    // kinetic denominator) instead of fourteen.  Same values as dividing in place, ~3x fewer instructions per call.
    struct Coef {
        double disp[4], film_f[4], film_s[4];
        double conv, inv_keq, Vf_Ks, R_AS, R_AW, R_Es, rho_B;
    };
    __host__ __device__ static Coef coefficients(const double* par) {
        Coef c;
        const double dx = par[0], eps = par[1];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            c.disp[i] = par[13 + i] / (dx * dx);
            c.film_f[i] = par[9 + i] / eps;
            c.film_s[i] = par[9 + i] / (1.0 - eps);
        }
        c.conv = par[2] / dx;
        c.inv_keq = 1.0 / par[7];
        c.Vf_Ks = par[4]; c.R_AS = par[5]; c.R_AW = par[6]; c.R_Es = par[8]; c.rho_B = par[3];
        return c;
    }
    __host__ __device__ static void rhs(const Coef& c, bool inlet, bool outlet, const double* cl, const double* cm,
                                        const double* cr, const double* cs, double* dcf, double* dcs) {
        if (inlet) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { dcf[i] = 0.0; dcs[i] = 0.0; }
            return;
        }
        const double v = div_z(div_z(c.Vf_Ks * (cs[0] * cs[1] - c.inv_keq * cs[2] * cs[3]),
                                     cs[1] + c.R_AS * cs[0] + c.R_AW * cs[3] + c.R_Es * cs[2]), 60.0);   // x / y / 60, zeros on the fast path (fdiv.cuh)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double right = outlet ? cl[i] : cr[i];
            dcf[i] = c.disp[i] * (cl[i] - 2.0 * cm[i] + right) - c.conv * (cm[i] - cl[i]) + c.film_f[i] * (cs[i] - cm[i]);
            const double sto = i < 2 ? -1.0 : 1.0;
            dcs[i] = -c.film_s[i] * (cs[i] - cm[i]) + c.rho_B * sto * v;
        }
    }
};

}  // namespace sodm
