// fpsb_krylov.cu — SpMV / fused two-column SpMM and the fused LSQR / CRAIG / MINRES / CGLS
// iterations of the IterativeSolver path.
//
// Reference surface replaced (file:line under /root/reference):
//   jac_op! products                         NLPModels, used at src/solve_linear_system.jl:119-121
//   solve_least_square  (LSQR on A')         src/solve_two_systems_struct.jl:167-185
//   solve_least_norm    (CRAIG, M=I/delta)   src/solve_two_systems_struct.jl:210-244
//   solve_two_mixed / least_squares / extras src/solve_linear_system.jl:45-140 (Iterative), :142-159 (LDLt extras)
//
// Design (B200): the Jacobian lives in HBM as CSR(A) and CSR(A') (no atomics, deterministic).
// One CTA per row block: the block's values+indices are staged into shared memory with two TMA
// 1-D bulk copies (cp.async.bulk + mbarrier), a sub-warp reduces every row against BOTH
// right-hand-side columns with a single 16-byte gather per nonzero (the two Golub-Kahan
// vectors are interleaved), and the row epilogue applies the axpby / norm / delayed x,w updates
// of the Krylov method so that every vector is read and written once per iteration.
// Norms use fixed-order reductions (per-CTA partials, last CTA finishes) and the scalar
// recurrences (Givens rotations, stopping tests) run on the device in the last CTA; the host only
// polls a done flag every few iterations.
#include "fpsb_internal.h"
#include "fpsb_device.cuh"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace fpsb {

static const double kSqrtEps = 1.4901161193847656e-08;

enum Mode {
    MD_NONE = 0,
    MD_PLAIN,
    MD_LSQR_INIT_M,   // v1 = A' u1 : first half step (row space of the LSQR solution)
    MD_LSQR_U,        // u <- Op v - alpha u          (row space of the LSQR right-hand side)
    MD_LSQR_V,        // w,x updates; v <- Op' u - beta v
    MD_CRAIG_V,       // delayed x,w2 updates; v <- Op' u - beta v
    MD_CRAIG_U,       // w,y updates; Mu <- Op v - alpha Mu
    MD_MINRES_M,
    MD_CGLS_INIT_M,
    MD_CGLS_N,
    MD_CGLS_M
};

enum EwOp {
    EW_COPY = 0,          // out = c0 * in
    EW_INIT_LSQR,         // self = b ; acc ||b||^2
    EW_INIT_CRAIG,        // self = c0 * b ; w = y = 0 ; acc
    EW_CRAIG_FLUSH,       // out = -(x + pending update)
    EW_MINRES_INIT,
    EW_MINRES_E1,
    EW_MINRES_E2,
    EW_CGLS_INIT,
    EW_CGLS_EN,
    EW_CGLS_EM
};

struct SlotIO {
    int mode;
    int pad;
    const double *gin;   // gather source (non-pair kernels)
    double *self;        // recurred vector of this row space (non-pair kernels)
    double *a0, *a1, *a2;
    double c0, c1;
};

struct StepParams {
    // SELL-32-sigma operator
    const int *wchunk;     // grid_sell * 8 + 1 : first slice of every warp's contiguous chunk
    const int *sl_off;     // nslice + 1 : element offsets (multiples of 32)
    const int *rowidx;     // nslice * 32 : original row of each lane, -1 for padding lanes
    const int *scol;
    const double *sval;
    int nslice;
    int grid_sell;         // CTAs working on slices; CTAs beyond handle one long row each
    // long rows (CSR)
    const int *long_row;
    const int *long_rp;
    const int *long_col;
    const double *long_val;
    int nrows;
    const double2 *gin2;   // interleaved gather pair (PAIR kernels)
    double2 *self2;        // interleaved recurred pair of this row space (PAIR kernels)
    SlotIO io[2];
    SlotState *st;
    double *partials;      // [grid][4]
    unsigned *counter;
    int *done_flag;        // set to 1 when no slot remains active
};

// ------------------------------------------------------------------------------------------------
// scalar recurrences (one thread, last CTA) — line-by-line the reference algorithms
// ------------------------------------------------------------------------------------------------
__device__ void slot_stop(SlotState &S) { S.active = 0; }

__device__ void lsqr_status(SlotState &S) {
    int st = FPSB_ST_UNKNOWN;
    if (S.tired) st = FPSB_ST_TIRED;
    if (S.ill_mach) st = FPSB_ST_ILLCOND_MACH;
    if (S.ill_lim) st = FPSB_ST_ILLCOND_LIM;
    if (S.solved) st = FPSB_ST_SOLVED;
    if (S.zero_resid) st = FPSB_ST_ZERO_RESID;
    if (S.fwd_err) st = FPSB_ST_FWD_ERR;
    S.status = st;
    S.inconsistent = !S.zero_resid;
}

__device__ void fin_init_lsqr(SlotState &S, double bb) {
    double beta1 = sqrt(bb);
    S.beta1 = beta1;
    S.iter = 0;
    if (beta1 == 0.0) {
        S.active = 0; S.solved = 1; S.inconsistent = 0; S.status = FPSB_ST_ZERO_RHS;
        return;
    }
    S.beta = beta1;
    S.su = 1.0 / beta1;
}

__device__ void fin_lsqr_init_m(SlotState &S, double vv) {
    S.Anorm2 = vv;
    S.Anorm = sqrt(vv);
    S.alpha = S.Anorm;
    S.Acond = 0; S.xNorm = 0; S.xNorm2 = 0; S.dNorm2 = 0;
    S.c2 = -1.0; S.s2 = 0.0; S.z = 0.0;
    S.xENorm2 = 0; S.err_lbnd = 0;
    for (int i = 0; i < 5; ++i) S.err_vec[i] = 0;
    S.rNorm = S.beta1; S.res2 = 0;
    S.ArNorm = S.ArNorm0 = S.alpha * S.beta;
    if (S.alpha == 0.0) {
        S.active = 0; S.solved = 1; S.inconsistent = 0; S.status = FPSB_ST_ZERO_ATB;
        return;
    }
    S.sv = 1.0 / S.alpha;
    S.phibar = S.beta1;
    S.rhobar = S.alpha;
    S.first = 1;
    int solved_lim = S.ArNorm / (S.Anorm * S.rNorm) <= S.axtol;
    int solved_mach = 1.0 + S.ArNorm / (S.Anorm * S.rNorm) <= 1.0;
    S.solved = solved_mach | solved_lim;
    S.tired = S.iter >= S.itmax;
    S.zero_resid = (S.rNorm / S.beta1 <= S.axtol) | (1.0 + S.rNorm / S.beta1 <= 1.0);
    S.ill_mach = S.ill_lim = S.fwd_err = 0;
    if (S.solved || S.tired) { lsqr_status(S); S.active = 0; }
}

// after u <- Op v - alpha u
__device__ void fin_lsqr_u(SlotState &S, double uu) {
    S.iter += 1;
    double beta = sqrt(uu);
    S.beta = beta;
    if (beta != 0.0) {
        S.su = 1.0 / beta;
        S.beta_zero = 0;
        S.Anorm2 = S.Anorm2 + S.alpha * S.alpha + beta * beta;
        if (S.lambda > 0) S.Anorm2 += S.lambda * S.lambda;
    } else {
        S.su = 0.0;
        S.beta_zero = 1;
    }
    double c1, s1, rhobar1;
    sym_givens(S.rhobar, S.lambda, c1, s1, rhobar1);
    S.psi = s1 * S.phibar;
    S.phibar = c1 * S.phibar;
    sym_givens(rhobar1, beta, S.c, S.s, S.rho);
    S.phi = S.c * S.phibar;
    S.phibar = S.s * S.phibar;
    S.xENorm2 += S.phi * S.phi;
    S.err_vec[S.iter % 5] = S.phi;
    if (S.iter >= 5) {
        double e = 0;
        for (int i = 0; i < 5; ++i) e += S.err_vec[i] * S.err_vec[i];
        S.err_lbnd = sqrt(e);
    }
    S.tau = S.s * S.phi;
    S.sigma = S.phi / S.rho;
}

// after v <- Op' u - beta v (and the x, w updates)
__device__ void fin_lsqr_v(SlotState &S, double vv, double ww) {
    if (!S.beta_zero) {
        S.alpha = sqrt(vv);
        S.sv = (S.alpha != 0.0) ? 1.0 / S.alpha : 0.0;
    }
    double theta = S.s * S.alpha;
    S.rhobar = -S.c * S.alpha;
    S.dNorm2 += ww / (S.rho * S.rho);
    S.tr_prev = theta / S.rho;
    S.first = 0;
    double delta = S.s2 * S.rho;
    double gammabar = -S.c2 * S.rho;
    double rhs = S.phi - delta * S.z;
    double zbar = rhs / gammabar;
    S.xNorm = sqrt(S.xNorm2 + zbar * zbar);
    double gamma;
    sym_givens(gammabar, theta, S.c2, S.s2, gamma);
    S.z = rhs / gamma;
    S.xNorm2 += S.z * S.z;
    S.Anorm = sqrt(S.Anorm2);
    S.Acond = S.Anorm * sqrt(S.dNorm2);
    double res1 = S.phibar * S.phibar;
    S.res2 += S.psi * S.psi;
    S.rNorm = sqrt(res1 + S.res2);
    S.ArNorm = S.alpha * fabs(S.tau);
    double test1 = S.rNorm / S.beta1;
    double test2 = S.ArNorm / (S.Anorm * S.rNorm);
    double test3 = 1.0 / S.Acond;
    double t1 = test1 / (1.0 + S.Anorm * S.xNorm / S.beta1);
    double rNormtol = S.btol + S.axtol * S.Anorm * S.xNorm / S.beta1;
    S.ill_mach = (1.0 + test3 <= 1.0);
    int solved_mach = (1.0 + test2 <= 1.0);
    int zero_resid_mach = (1.0 + t1 <= 1.0);
    S.tired = S.iter >= S.itmax;
    S.ill_lim = (test3 <= S.ctol);
    int solved_lim = (test2 <= S.axtol);
    int solved_opt = S.ArNorm <= S.atol + S.rtol * S.ArNorm0;
    int zero_resid_lim = (test1 <= rNormtol);
    if (S.iter >= 5) S.fwd_err = S.err_lbnd <= S.etol * sqrt(S.xENorm2);
    int ill_cond = S.ill_mach || S.ill_lim;
    S.zero_resid = zero_resid_mach || zero_resid_lim;
    S.solved = solved_mach || solved_lim || solved_opt || S.zero_resid || S.fwd_err;
    if (S.solved || S.tired || ill_cond) { lsqr_status(S); S.active = 0; }
}

__device__ void craig_status(SlotState &S) {
    int st = FPSB_ST_UNKNOWN;
    if (S.tired) st = FPSB_ST_TIRED;
    if (S.solved) st = FPSB_ST_SOLVED;
    if (S.ill_mach) st = FPSB_ST_ILLCOND_MACH;
    if (S.ill_lim) st = FPSB_ST_ILLCOND_LIM;
    if (S.inconsistent) st = FPSB_ST_INCONSISTENT;
    S.status = st;
}

__device__ void fin_init_craig(SlotState &S, double bb) {
    double beta1 = sqrt(S.mscale * bb);   // sqrt(u'Mu), u = mscale * Mu
    S.beta1 = beta1;
    S.rNorm = beta1;
    S.iter = 0;
    if (beta1 == 0.0) {
        S.active = 0; S.solved = 1; S.inconsistent = 0; S.status = FPSB_ST_ZERO_RHS;
        return;
    }
    S.beta1sq = beta1 * beta1;
    S.beta = beta1;
    S.theta = beta1;
    S.xi = -1.0;
    S.deltag = S.lambda;
    S.rho_prev = 1.0;
    S.su = 1.0 / beta1;
    S.sv = 0.0;
    S.pend = 0;
    S.c1 = 1.0; S.s1 = 0.0; S.s2g = 1.0;
    S.Anorm2 = 0; S.Anorm = 0; S.dNorm2 = 0; S.Acond = 0; S.xNorm2 = 0;
    S.eps_c = S.atol + S.rtol * S.rNorm;
    double bkwerr = 1.0;
    int solved_lim = bkwerr <= S.btol;
    int solved_mach = 1.0 + bkwerr <= 1.0;
    int solved_resid_tol = S.rNorm <= S.eps_c;
    int solved_resid_lim = S.rNorm <= S.btol + S.atol * S.Anorm * sqrt(S.xNorm2) / beta1;
    S.solved = solved_mach | solved_lim | solved_resid_tol | solved_resid_lim;
    S.tired = S.iter >= S.itmax;
    S.ill_mach = S.ill_lim = 0;
    S.inconsistent = 0;
    if (S.solved || S.tired) { craig_status(S); S.active = 0; }
}

// after v <- Op' u - beta v
__device__ void fin_craig_v(SlotState &S, double vv) {
    S.pend = 0;   // the previous iteration's x update was applied by this kernel
    double alpha = sqrt(vv);
    if (alpha == 0.0) {
        S.inconsistent = 1;
        craig_status(S);
        S.active = 0;
        return;
    }
    S.alpha = alpha;
    S.sv = 1.0 / alpha;
    S.Anorm2 += alpha * alpha;
    if (S.lambda > 0) sym_givens(alpha, S.deltag, S.c1, S.s1, S.rho);
    else S.rho = alpha;
    S.xi = -S.theta / S.rho * S.xi;
    S.trw = S.theta / S.rho_prev;
    S.xr = S.xi / S.rho;
    S.pend = 1;
}

// after Mu <- Op v - alpha Mu (and the w, y updates)
__device__ void fin_craig_u(SlotState &S, double uu, double ww) {
    double beta = sqrt(S.mscale * uu);
    S.beta = beta;
    S.su = (beta != 0.0) ? 1.0 / beta : 0.0;
    if (S.lambda > 0) {
        S.theta = S.c1 * beta;
        double gamma = S.s1 * beta;
        double c2;
        sym_givens(S.lambda, gamma, c2, S.s2g, S.deltag);
    } else {
        S.theta = beta;
    }
    S.Anorm2 += beta * beta;
    S.Anorm = sqrt(S.Anorm2);
    S.dNorm2 += sqrt(ww);   // upstream craig.jl accumulates the 2-norm here (kept)
    S.Acond = S.Anorm * sqrt(S.dNorm2);
    S.xNorm2 += S.xi * S.xi;
    S.rNorm = beta * fabs(S.xi);
    if (S.lambda > 0) S.rNorm *= fabs(S.c1);
    S.iter += 1;
    double bkwerr = S.rNorm / sqrt(S.beta1sq + S.Anorm2 * S.xNorm2);
    S.rho_prev = S.rho;
    int solved_lim = bkwerr <= S.btol;
    int solved_mach = 1.0 + bkwerr <= 1.0;
    int solved_resid_tol = S.rNorm <= S.eps_c;
    int solved_resid_lim = S.rNorm <= S.btol + S.atol * S.Anorm * sqrt(S.xNorm2) / S.beta1;
    S.solved = solved_mach | solved_lim | solved_resid_tol | solved_resid_lim;
    S.ill_mach = 1.0 + 1.0 / S.Acond <= 1.0;
    S.ill_lim = 1.0 / S.Acond <= S.ctol;
    S.inconsistent = 0;
    S.tired = S.iter >= S.itmax;
    S.xNorm = sqrt(S.xNorm2);
    if (S.solved || S.ill_mach || S.ill_lim || S.tired) { craig_status(S); S.active = 0; }
}

// ---- MINRES ------------------------------------------------------------------------------------
__device__ void fin_minres_init(SlotState &S, double bb) {
    S.iter = 0;
    if (bb == 0.0) {
        S.active = 0; S.solved = 1; S.inconsistent = 0; S.status = FPSB_ST_ZERO_RHS;
        S.beta1 = 0;
        return;
    }
    double beta1 = sqrt(bb);
    S.beta1 = beta1; S.beta = beta1; S.oldbeta = 0; S.deltabar = 0; S.eps_ = 0;
    S.rNorm = beta1; S.phibar = beta1; S.rhs1 = beta1; S.rhs2 = 0;
    S.gmax = 0; S.gmin = INFINITY; S.cs = -1.0; S.sn = 0;
    S.Anorm2 = 0; S.Anorm = 0; S.Acond = 0; S.ArNorm = 0; S.xNorm = 0; S.xENorm2 = 0; S.err_lbnd = 0;
    for (int i = 0; i < 5; ++i) S.err_vec[i] = 0;
    S.tol = S.atol + S.rtol * beta1;
    S.solved = (S.rNorm <= S.rtol);
    S.tired = S.iter >= S.itmax;
    S.zero_resid = (S.rNorm <= S.tol);
    S.ill_mach = S.ill_lim = S.fwd_err = 0;
    if (S.solved || S.tired) {
        S.status = S.solved ? FPSB_ST_SOLVED : FPSB_ST_TIRED;
        S.inconsistent = !S.zero_resid;
        S.active = 0;
    }
}
__device__ void fin_minres_m(SlotState &S, double vy) {
    S.iter += 1;
    S.alpha = vy / S.beta;
    S.delta = S.cs * S.deltabar + S.sn * S.alpha;
}
__device__ void fin_minres_e1(SlotState &S, double yy) {
    S.oldbeta = S.beta;
    S.beta = sqrt(yy);
    double alpha = S.alpha, beta = S.beta;
    S.Anorm2 = S.Anorm2 + alpha * alpha + S.oldbeta * S.oldbeta + beta * beta;
    S.gammabar = S.sn * S.deltabar - S.cs * alpha;
    S.eps_ = S.sn * beta;
    S.deltabar = -S.cs * beta;
    S.root = sqrt(S.gammabar * S.gammabar + S.deltabar * S.deltabar);
    S.ArNorm = S.phibar * S.root;
    double gamma = sqrt(S.gammabar * S.gammabar + beta * beta);
    gamma = fmax(gamma, 2.220446049250313e-16);
    S.gamma = gamma;
    S.cs = S.gammabar / gamma;
    S.sn = beta / gamma;
    S.phi = S.cs * S.phibar;
    S.phibar = S.sn * S.phibar;
}
__device__ void fin_minres_e2(SlotState &S, double xx) {
    const double epsM = 2.220446049250313e-16;
    S.xENorm2 += S.phi * S.phi;
    S.err_vec[S.iter % 5] = S.phi;
    if (S.iter >= 5) {
        double e = 0;
        for (int i = 0; i < 5; ++i) e += S.err_vec[i] * S.err_vec[i];
        S.err_lbnd = sqrt(e);
    }
    S.gmax = fmax(S.gmax, S.gamma);
    S.gmin = fmin(S.gmin, S.gamma);
    double zeta = S.rhs1 / S.gamma;
    S.rhs1 = S.rhs2 - S.delta * zeta;
    S.rhs2 = -S.eps_ * zeta;
    S.Anorm = sqrt(S.Anorm2);
    S.xNorm = sqrt(xx);
    S.rNorm = S.phibar;
    double test1 = S.rNorm / (S.Anorm * S.xNorm);
    double test2 = S.root / S.Anorm;
    S.Acond = S.gmax / S.gmin;
    if (S.iter == 1 && S.beta / S.beta1 <= 10 * epsM) {
        S.solved = 1; S.inconsistent = 1; S.status = FPSB_ST_ZERO_ATB; S.active = 0;
        return;
    }
    S.ill_mach = (1.0 + 1.0 / S.Acond <= 1.0);
    int solved_mach = (1.0 + test2 <= 1.0);
    int zero_resid_mach = (1.0 + test1 <= 1.0);
    int resid_decrease_mach = (S.rNorm + 1.0 <= 1.0);
    S.tired = S.iter >= S.itmax;
    S.ill_lim = (1.0 / S.Acond <= S.ctol);
    int solved_lim = (test2 <= S.tol);
    int zero_resid_lim = (test1 <= S.tol);
    int resid_decrease_lim = (S.rNorm <= S.tol);
    if (S.iter >= 5) S.fwd_err = S.err_lbnd <= S.etol * sqrt(S.xENorm2);
    S.zero_resid = zero_resid_mach | zero_resid_lim;
    int resid_decrease = resid_decrease_mach | resid_decrease_lim;
    int ill_cond = S.ill_mach | S.ill_lim;
    S.solved = solved_mach | solved_lim | S.zero_resid | S.fwd_err | resid_decrease;
    if (S.solved || S.tired || ill_cond) {
        int st = FPSB_ST_UNKNOWN;
        if (S.tired) st = FPSB_ST_TIRED;
        if (S.ill_mach) st = FPSB_ST_ILLCOND_MACH;
        if (S.ill_lim) st = FPSB_ST_ILLCOND_LIM;
        if (S.solved) st = FPSB_ST_SOLVED;
        if (S.zero_resid) st = FPSB_ST_ZERO_RESID;
        if (S.fwd_err) st = FPSB_ST_FWD_ERR;
        if (resid_decrease) st = FPSB_ST_SOLVED;
        S.status = st;
        S.inconsistent = !S.zero_resid;
        S.active = 0;
    }
}

// ---- CGLS --------------------------------------------------------------------------------------
__device__ void fin_cgls_init(SlotState &S, double bb) {
    S.iter = 0;
    S.bnorm = sqrt(bb);
    S.rNorm = S.bnorm;
    if (S.bnorm == 0.0) { S.active = 0; S.solved = 1; S.status = FPSB_ST_ZERO_RHS; }
}
__device__ void fin_cgls_init_m(SlotState &S, double ss) {
    S.gamma_c = ss;
    S.pp = ss;
    S.ArNorm = sqrt(ss);
    S.tol = S.atol + S.rtol * S.ArNorm;
    S.solved = S.ArNorm <= S.tol;
    S.tired = S.iter >= S.itmax;
    if (S.solved || S.tired) { S.status = S.solved ? FPSB_ST_SOLVED : FPSB_ST_TIRED; S.active = 0; }
}
__device__ void fin_cgls_n(SlotState &S, double qq) {
    double delta = qq;
    if (S.lambda > 0) delta += S.lambda * S.pp;
    S.alpha = S.gamma_c / delta;
}
__device__ void fin_cgls_en(SlotState &S, double rr) { S.rNorm = sqrt(rr); }
__device__ void fin_cgls_m(SlotState &S, double ss) {
    S.beta_c = ss / S.gamma_c;
    S.gamma_c = ss;
}
__device__ void fin_cgls_em(SlotState &S, double pp) {
    S.pp = pp;
    S.ArNorm = sqrt(S.gamma_c);
    S.iter += 1;
    S.solved = S.ArNorm <= S.tol;
    S.tired = S.iter >= S.itmax;
    if (S.solved || S.tired) { S.status = S.solved ? FPSB_ST_SOLVED : FPSB_ST_TIRED; S.active = 0; }
}

__device__ void finish_step(SlotState &S, int mode, double a0, double a1) {
    switch (mode) {
        case MD_LSQR_INIT_M: fin_lsqr_init_m(S, a0); break;
        case MD_LSQR_U: fin_lsqr_u(S, a0); break;
        case MD_LSQR_V: fin_lsqr_v(S, a0, a1); break;
        case MD_CRAIG_V: fin_craig_v(S, a0); break;
        case MD_CRAIG_U: fin_craig_u(S, a0, a1); break;
        case MD_MINRES_M: fin_minres_m(S, a0); break;
        case MD_CGLS_INIT_M: fin_cgls_init_m(S, a0); break;
        case MD_CGLS_N: fin_cgls_n(S, a0); break;
        case MD_CGLS_M: fin_cgls_m(S, a0); break;
        default: break;
    }
}
__device__ void finish_ew(SlotState &S, int op, double a0) {
    switch (op) {
        case EW_INIT_LSQR: fin_init_lsqr(S, a0); break;
        case EW_INIT_CRAIG: fin_init_craig(S, a0); break;
        case EW_MINRES_INIT: fin_minres_init(S, a0); break;
        case EW_MINRES_E1: fin_minres_e1(S, a0); break;
        case EW_MINRES_E2: fin_minres_e2(S, a0); break;
        case EW_CGLS_INIT: fin_cgls_init(S, a0); break;
        case EW_CGLS_EN: fin_cgls_en(S, a0); break;
        case EW_CGLS_EM: fin_cgls_em(S, a0); break;
        default: break;
    }
}

// ------------------------------------------------------------------------------------------------
// per-slot coefficients loaded once per CTA
// ------------------------------------------------------------------------------------------------
struct Coef {
    int mode, first, pend, iter;
    int rd0, rd1, wr0, wr1;      // which aux vectors (a0 / a1) the row epilogue reads / writes
    int rdself;
    double gsc, ssc, alpha, beta, sigma, tr_prev, lambda;
    double xi, c1, s1, s2g, trw, xr, mscale, oldbeta, c0, c1h;
};

__device__ __forceinline__ void load_coef(Coef &C, const SlotIO &io, const SlotState *st, bool use_state) {
    C.mode = io.mode;
    C.c0 = io.c0; C.c1h = io.c1;
    C.gsc = 1.0; C.ssc = 1.0;
    C.rd0 = C.rd1 = C.wr0 = C.wr1 = C.rdself = 0;
    C.first = 0; C.pend = 0; C.iter = 0; C.lambda = 0.0;
    if (io.mode == MD_NONE) return;
    if (io.mode == MD_PLAIN) { C.rd0 = io.a0 != nullptr; return; }
    if (!use_state) return;
    C.first = st->first; C.pend = st->pend; C.iter = st->iter;
    C.alpha = st->alpha; C.beta = st->beta; C.sigma = st->sigma; C.tr_prev = st->tr_prev;
    C.lambda = st->lambda; C.xi = st->xi; C.c1 = st->c1; C.s1 = st->s1; C.s2g = st->s2g;
    C.trw = st->trw; C.xr = st->xr; C.mscale = st->mscale; C.oldbeta = st->oldbeta;
    switch (io.mode) {
        case MD_LSQR_INIT_M: C.gsc = st->su; break;
        case MD_LSQR_U: C.gsc = st->sv; C.ssc = st->su; C.rdself = 1; break;
        case MD_LSQR_V:
            C.gsc = st->su; C.ssc = st->sv; C.rdself = 1;
            C.rd0 = C.rd1 = !C.first; C.wr0 = C.wr1 = 1;
            break;
        case MD_CRAIG_V:
            C.gsc = st->mscale * st->su; C.ssc = st->sv; C.rdself = 1;
            C.rd0 = C.wr0 = C.pend;
            C.rd1 = C.wr1 = C.pend && (C.lambda > 0);
            break;
        case MD_CRAIG_U:
            C.gsc = st->sv; C.ssc = st->su; C.rdself = 1;
            C.rd0 = C.rd1 = C.wr0 = C.wr1 = 1;
            break;
        case MD_MINRES_M: C.rd0 = 1; C.rd1 = (C.iter + 1 >= 2); break;
        case MD_CGLS_M: C.rd0 = C.rd1 = 1; C.wr0 = 1; break;
        default: break;
    }
}

// row epilogue on values: (sraw, selfold, a0, a1) -> returns new self; a0 / a1 updated in place
__device__ __forceinline__ double row_epilogue(const Coef &C, double sraw, double selfold, double &a0,
                                               double &a1, double &acc0, double &acc1) {
    double out = 0.0;
    switch (C.mode) {
        case MD_PLAIN: {
            out = C.c0 * sraw;
            if (C.rd0) out += C.c1h * a0;
            acc0 += out * out;
        } break;
        case MD_LSQR_INIT_M: {
            out = C.gsc * sraw;
            acc0 += out * out;
        } break;
        case MD_LSQR_U: {
            out = C.gsc * sraw - C.alpha * (selfold * C.ssc);
            acc0 += out * out;
        } break;
        case MD_LSQR_V: {
            const double vj = selfold * C.ssc;
            const double wj = C.first ? vj : (vj - C.tr_prev * a0);
            acc1 += wj * wj;
            a0 = wj;
            a1 = (C.first ? 0.0 : a1) + C.sigma * wj;
            out = C.gsc * sraw - C.beta * vj;
            acc0 += out * out;
        } break;
        case MD_CRAIG_V: {
            const double vp = selfold * C.ssc;
            if (C.pend) {
                if (C.lambda > 0) {
                    const double w2 = a1;
                    double x = a0 + (C.xi * C.c1) * vp;
                    x = x + (C.xi * C.s1) * w2;
                    a0 = x;
                    a1 = C.s2g * (C.s1 * vp - C.c1 * w2);
                } else {
                    a0 += C.xi * vp;
                }
            }
            out = C.gsc * sraw - C.beta * vp;
            acc0 += out * out;
        } break;
        case MD_CRAIG_U: {
            const double mu = selfold * C.ssc;
            const double uj = C.mscale * mu;
            const double wj = uj - C.trw * a0;
            a0 = wj;
            a1 += C.xr * wj;
            acc1 += wj * wj;
            out = C.gsc * sraw - C.alpha * mu;
            acc0 += out * out;
        } break;
        case MD_MINRES_M: {
            // a0 = r2 (= v), a1 = r1 ; self = y
            const double r2 = a0;
            double y = sraw;
            if (C.lambda != 0.0) y += C.lambda * r2;
            y *= (1.0 / C.beta);
            if (C.rd1) y -= (C.beta / C.oldbeta) * a1;
            out = y;
            acc0 += r2 * y;
        } break;
        case MD_CGLS_INIT_M: {
            out = sraw;           // p = s
            acc0 += out * out;
        } break;
        case MD_CGLS_N: {
            out = sraw;           // q
            acc0 += out * out;
        } break;
        case MD_CGLS_M: {
            // a0 = x, a1 = p ; self = s
            const double x = a0 + C.alpha * a1;
            a0 = x;
            double sv = sraw;
            if (C.lambda > 0) sv -= C.lambda * x;
            out = sv;
            acc0 += sv * sv;
        } break;
        default: break;
    }
    return out;
}

// ------------------------------------------------------------------------------------------------
// the fused SpMM step kernel — SELL-32-sigma streamed through a TMA ring.
//
// Layout: a slice is 32 rows, stored column-major (32 consecutive entries per slice column), lane ==
// row.  A CTA owns a contiguous range of slices, i.e. one contiguous stream of 32-entry "stream
// rows" of sval / scol.  Warp 8 is the producer: one lane moves the stream into an 8-stage ring in
// shared memory with cp.async.bulk (TMA), 32 stream rows (8 KB of values + 4 KB of indices) per
// stage, signalled by mbarriers — 96 KB of loads in flight per CTA, none of it held in registers.
// Warps 0-7 are consumers: they take the CTA's slices round-robin (adjacent slices => the 8 gather
// windows overlap in L1), read values / indices conflict-free from the ring, keep only the
// x-gathers in flight, accumulate the two row sums in registers and run the Krylov row epilogue in
// the same lane (operands prefetched at slice start).  Every consumer walks every stage (waits
// full, arrives empty), so a stage is recycled exactly when all 8 warps are past it.
// Rows longer than kLongRow are handled by extra CTAs (one per row, block reduction).
// ------------------------------------------------------------------------------------------------
constexpr int kUnroll = 4;
constexpr int kConsumers = 8;
constexpr int kStepThreads = (kConsumers + 1) * 32;
constexpr int kRingStages = 8;
constexpr int kStageRows = 32;
constexpr int kStageElems = kStageRows * 32;
#ifdef FPSB_STEP_PLAIN
constexpr int kRingBytes = 0;
#else
constexpr int kRingBytes = kRingStages * kStageElems * 12;
#endif
constexpr int kMaxSlicesPerCta = 96;     // slice metadata staged in shared memory (12.4 KB)
static_assert(kStageRows == 32 && kStageElems == 1024 && (kRingStages & (kRingStages - 1)) == 0, "ring indexing uses shifts");

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <bool PAIR>
__global__ void __launch_bounds__(kStepThreads, 2) gk_step_kernel(StepParams P, int use_state) {
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ double s_red[4 * 16];
    __shared__ int s_off[kMaxSlicesPerCta + 1];
    __shared__ int s_row[kMaxSlicesPerCta * 32];
    __shared__ alignas(8) uint64_t full_bar[kRingStages];
    __shared__ alignas(8) uint64_t empty_bar[kRingStages];
    __shared__ int s_last;

    const int tid = threadIdx.x;
    const bool act0 = P.io[0].mode != MD_NONE && (!use_state || P.st[0].active);
    const bool act1 = P.io[1].mode != MD_NONE && (!use_state || P.st[1].active);
    if (!act0 && !act1) return;

    // per-slot coefficients live in shared memory (broadcast reads) to keep registers for the
    // loads in flight
    __shared__ Coef sC[2];
    if (tid == 0) {
        load_coef(sC[0], P.io[0], &P.st[0], use_state);
        load_coef(sC[1], P.io[1], &P.st[1], use_state);
        if (!act0) { sC[0].mode = MD_NONE; sC[0].rd0 = sC[0].rd1 = sC[0].wr0 = sC[0].wr1 = sC[0].rdself = 0; }
        if (!act1) { sC[1].mode = MD_NONE; sC[1].rd0 = sC[1].rd1 = sC[1].wr0 = sC[1].wr1 = sC[1].rdself = 0; }
        for (int s = 0; s < kRingStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], kConsumers); }
        mbar_fence_init();
    }
    __syncthreads();
    const Coef &C0 = sC[0];
    const Coef &C1 = sC[1];

    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    const int cta = blockIdx.x;

    if (cta < P.grid_sell) {
        const int lane = tid & 31, wid = tid >> 5;
        const int s_begin = (int)(((int64_t)P.nslice * cta) / P.grid_sell);
        const int s_end = (int)(((int64_t)P.nslice * (cta + 1)) / P.grid_sell);
        // stage this CTA's slice metadata (row offsets and the lane -> row map) in shared memory
        for (int i = tid; i <= s_end - s_begin; i += kStepThreads) s_off[i] = P.sl_off[s_begin + i] >> 5;
        for (int i = tid; i < (s_end - s_begin) * 32; i += kStepThreads) s_row[i] = P.rowidx[s_begin * 32 + i];
        __syncthreads();
        const int R0 = s_off[0], R1 = s_off[s_end - s_begin];
        const int nstage = (R1 - R0 + kStageRows - 1) / kStageRows;
        const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
        double *ring_val = reinterpret_cast<double *>(ring);
        int *ring_col = reinterpret_cast<int *>(ring + (size_t)kRingStages * kStageElems * 8);
        bool ok = true;
#ifdef FPSB_STEP_PLAIN
        // A/B variant: no TMA ring — all 9 warps stream their slices with plain coalesced loads
        // (two register batches in flight), same L2 policies, staged metadata and operand prefetch
        {
            (void)pol_stream; (void)ring_val; (void)ring_col; (void)nstage;
            struct Ops { double2 old2; double so0, so1, a00, a01, a10, a11; };
            auto load_ops = [&](int row, Ops &o) {
                o.old2 = make_double2(0.0, 0.0);
                o.so0 = o.so1 = o.a00 = o.a01 = o.a10 = o.a11 = 0.0;
                if (row >= 0) {
                    if (PAIR) o.old2 = __ldcs(P.self2 + row);
                    else {
                        if (C0.rdself) o.so0 = __ldcs(P.io[0].self + row);
                        if (C1.rdself) o.so1 = __ldcs(P.io[1].self + row);
                    }
                    if (C0.rd0) o.a00 = __ldcs(P.io[0].a0 + row);
                    if (C0.rd1) o.a01 = __ldcs(P.io[0].a1 + row);
                    if (C1.rd0) o.a10 = __ldcs(P.io[1].a0 + row);
                    if (C1.rd1) o.a11 = __ldcs(P.io[1].a1 + row);
                }
            };
            constexpr int NW = kConsumers + 1;
            const int ns = s_end - s_begin;
            Ops nxt;
            if (wid < ns) load_ops(s_row[wid * 32 + lane], nxt);
            for (int ls = wid; ls < ns; ls += NW) {
                const int width = s_off[ls + 1] - s_off[ls];
                const int row = s_row[ls * 32 + lane];
                Ops o = nxt;
                if (ls + NW < ns) load_ops(s_row[(ls + NW) * 32 + lane], nxt);
                const double *vp = P.sval + (size_t)s_off[ls] * 32 + lane;
                const int *cp = P.scol + (size_t)s_off[ls] * 32 + lane;
                double s0 = 0.0, s1 = 0.0;
                const int nb = (width + kUnroll - 1) / kUnroll;
                double vA[kUnroll], vB[kUnroll];
                int cA[kUnroll], cB[kUnroll];
                auto load_batch = [&](int bidx, double (&v)[kUnroll], int (&c)[kUnroll]) {
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u) {
                        const int j = bidx * kUnroll + u;
                        const bool in = j < width;
                        v[u] = in ? __ldcs(vp + (size_t)j * 32) : 0.0;
                        c[u] = in ? __ldcs(cp + (size_t)j * 32) : -1;
                    }
                };
                auto consume = [&](const double (&v)[kUnroll], const int (&c)[kUnroll]) {
                    if (PAIR) {
                        double2 x[kUnroll];
#pragma unroll
                        for (int u = 0; u < kUnroll; ++u)
                            x[u] = (c[u] >= 0) ? ldg_evict_last(P.gin2 + c[u], pol_keep) : make_double2(0.0, 0.0);
#pragma unroll
                        for (int u = 0; u < kUnroll; ++u) { s0 = fma(v[u], x[u].x, s0); s1 = fma(v[u], x[u].y, s1); }
                    } else {
                        double x0[kUnroll], x1[kUnroll];
#pragma unroll
                        for (int u = 0; u < kUnroll; ++u) {
                            x0[u] = (act0 && c[u] >= 0) ? ldg_evict_last(P.io[0].gin + c[u], pol_keep) : 0.0;
                            x1[u] = (act1 && c[u] >= 0) ? ldg_evict_last(P.io[1].gin + c[u], pol_keep) : 0.0;
                        }
#pragma unroll
                        for (int u = 0; u < kUnroll; ++u) { s0 = fma(v[u], x0[u], s0); s1 = fma(v[u], x1[u], s1); }
                    }
                };
                load_batch(0, vA, cA);
                load_batch(1, vB, cB);
                for (int bb = 0; bb < nb; bb += 2) {
                    consume(vA, cA);
                    load_batch(bb + 2, vA, cA);
                    if (bb + 1 < nb) { consume(vB, cB); load_batch(bb + 3, vB, cB); }
                }
                if (row >= 0) {
                    if (PAIR) {
                        double2 nw = o.old2;
                        if (act0) nw.x = row_epilogue(C0, s0, o.old2.x, o.a00, o.a01, acc[0], acc[1]);
                        if (act1) nw.y = row_epilogue(C1, s1, o.old2.y, o.a10, o.a11, acc[2], acc[3]);
                        stg_evict_last(P.self2 + row, nw, pol_keep);
                    } else {
                        if (act0) __stcs(P.io[0].self + row, row_epilogue(C0, s0, o.so0, o.a00, o.a01, acc[0], acc[1]));
                        if (act1) __stcs(P.io[1].self + row, row_epilogue(C1, s1, o.so1, o.a10, o.a11, acc[2], acc[3]));
                    }
                    if (C0.wr0) __stcs(P.io[0].a0 + row, o.a00);
                    if (C0.wr1) __stcs(P.io[0].a1 + row, o.a01);
                    if (C1.wr0) __stcs(P.io[1].a0 + row, o.a10);
                    if (C1.wr1) __stcs(P.io[1].a1 + row, o.a11);
                }
            }
        }
#else
        if (wid == kConsumers) {
            // ---------------- producer ----------------
            if (lane == 0) {
                for (int k = 0; k < nstage && ok; ++k) {
                    const int slot = k % kRingStages;
                    if (k >= kRingStages) ok = mbar_wait(&empty_bar[slot], (uint32_t)(((k / kRingStages) & 1) ^ 1));
                    if (!ok) break;
                    const int rows = min(kStageRows, R1 - R0 - k * kStageRows);
                    const size_t e = ((size_t)R0 + (size_t)k * kStageRows) * 32;
                    mbar_expect_tx(&full_bar[slot], (uint32_t)rows * 32u * 12u);
                    tma_bulk_g2s_hint(ring_val + (size_t)slot * kStageElems, P.sval + e, (uint32_t)rows * 256u, &full_bar[slot], pol_stream);
                    tma_bulk_g2s_hint(ring_col + (size_t)slot * kStageElems, P.scol + e, (uint32_t)rows * 128u, &full_bar[slot], pol_stream);
                }
            }
        } else {
            // ---------------- consumers ----------------
            int kcur = -1;
            auto advance_to = [&](int k) {
                while (kcur < k && ok) {
                    if (kcur >= 0) { __syncwarp(); if (lane == 0) mbar_arrive(&empty_bar[kcur % kRingStages]); }
                    ++kcur;
                    ok = mbar_wait(&full_bar[kcur % kRingStages], (uint32_t)((kcur / kRingStages) & 1));
                }
            };
            // row-epilogue operands of a slice (registers); loaded one slice ahead so that neither
            // the row map nor these DRAM reads sit on the slice's critical path
            struct Ops { double2 old2; double so0, so1, a00, a01, a10, a11; };
            auto load_ops = [&](int row, Ops &o) {
                o.old2 = make_double2(0.0, 0.0);
                o.so0 = o.so1 = o.a00 = o.a01 = o.a10 = o.a11 = 0.0;
                if (row >= 0) {
                    if (PAIR) o.old2 = __ldcs(P.self2 + row);
                    else {
                        if (C0.rdself) o.so0 = __ldcs(P.io[0].self + row);
                        if (C1.rdself) o.so1 = __ldcs(P.io[1].self + row);
                    }
                    if (C0.rd0) o.a00 = __ldcs(P.io[0].a0 + row);
                    if (C0.rd1) o.a01 = __ldcs(P.io[0].a1 + row);
                    if (C1.rd0) o.a10 = __ldcs(P.io[1].a0 + row);
                    if (C1.rd1) o.a11 = __ldcs(P.io[1].a1 + row);
                }
            };
            const int ns = s_end - s_begin;
            Ops nxt;
            if (wid < ns) load_ops(s_row[wid * 32 + lane], nxt);
            for (int ls = wid; ls < ns && ok; ls += kConsumers) {
                const int ra = s_off[ls] - R0, rb = s_off[ls + 1] - R0;
                const int row = s_row[ls * 32 + lane];
                Ops o = nxt;
                if (ls + kConsumers < ns) load_ops(s_row[(ls + kConsumers) * 32 + lane], nxt);
                double s0 = 0.0, s1 = 0.0;
                // the slice's rows, one ring stage at a time (a 20-row slice touches 1-2 stages)
                int r = ra;
                while (r < rb && ok) {
                    const int k = r >> 5;                               // kStageRows == 32
                    advance_to(k);
                    const int seg_end = min(rb, (k + 1) << 5);
                    const int nrows = seg_end - r;
                    const int q0 = ((k & (kRingStages - 1)) << 10) + ((r & 31) << 5) + lane;
                    const double *vp = ring_val + q0;
                    const int *cp = ring_col + q0;
                    for (int j = 0; j < nrows; j += kUnroll) {
                        double v[kUnroll];
                        int c[kUnroll];
#pragma unroll
                        for (int u = 0; u < kUnroll; ++u) {
                            const bool in = (j + u) < nrows;
                            v[u] = in ? vp[(j + u) * 32] : 0.0;
                            c[u] = in ? cp[(j + u) * 32] : -1;
                        }
                        if (PAIR) {
                            double2 x[kUnroll];
#pragma unroll
                            for (int u = 0; u < kUnroll; ++u)
                                x[u] = (c[u] >= 0) ? ldg_evict_last(P.gin2 + c[u], pol_keep) : make_double2(0.0, 0.0);
#pragma unroll
                            for (int u = 0; u < kUnroll; ++u) { s0 = fma(v[u], x[u].x, s0); s1 = fma(v[u], x[u].y, s1); }
                        } else {
                            double x0[kUnroll], x1[kUnroll];
#pragma unroll
                            for (int u = 0; u < kUnroll; ++u) {
                                x0[u] = (act0 && c[u] >= 0) ? ldg_evict_last(P.io[0].gin + c[u], pol_keep) : 0.0;
                                x1[u] = (act1 && c[u] >= 0) ? ldg_evict_last(P.io[1].gin + c[u], pol_keep) : 0.0;
                            }
#pragma unroll
                            for (int u = 0; u < kUnroll; ++u) { s0 = fma(v[u], x0[u], s0); s1 = fma(v[u], x1[u], s1); }
                        }
                    }
                    r = seg_end;
                }
                if (row >= 0 && ok) {
                    if (PAIR) {
                        double2 nw = o.old2;
                        if (act0) nw.x = row_epilogue(C0, s0, o.old2.x, o.a00, o.a01, acc[0], acc[1]);
                        if (act1) nw.y = row_epilogue(C1, s1, o.old2.y, o.a10, o.a11, acc[2], acc[3]);
                        stg_evict_last(P.self2 + row, nw, pol_keep);      // gathered by the next launch
                    } else {
                        if (act0) __stcs(P.io[0].self + row, row_epilogue(C0, s0, o.so0, o.a00, o.a01, acc[0], acc[1]));
                        if (act1) __stcs(P.io[1].self + row, row_epilogue(C1, s1, o.so1, o.a10, o.a11, acc[2], acc[3]));
                    }
                    if (C0.wr0) __stcs(P.io[0].a0 + row, o.a00);
                    if (C0.wr1) __stcs(P.io[0].a1 + row, o.a01);
                    if (C1.wr0) __stcs(P.io[1].a0 + row, o.a10);
                    if (C1.wr1) __stcs(P.io[1].a1 + row, o.a11);
                }
            }
            // walk (and release) the remaining stages so the producer can finish
            advance_to(nstage - 1);
            if (kcur >= 0 && ok) { __syncwarp(); if (lane == 0) mbar_arrive(&empty_bar[kcur % kRingStages]); }
        }
#endif
        if (!ok && lane == 0) atomicExch(P.done_flag, -1);
    } else {
        // one long row per CTA: strided over the whole block, fixed-tree block reduction
        const int lr = cta - P.grid_sell;
        const int row = P.long_row[lr];
        const int e0 = P.long_rp[lr], e1 = P.long_rp[lr + 1];
        double a[2] = {0.0, 0.0};
        for (int k = e0 + tid; k < e1; k += kStepThreads) {
            const double v = P.long_val[k];
            const int c = P.long_col[k];
            if (PAIR) {
                const double2 x = __ldg(P.gin2 + c);
                a[0] += v * x.x; a[1] += v * x.y;
            } else {
                if (act0) a[0] += v * __ldg(P.io[0].gin + c);
                if (act1) a[1] += v * __ldg(P.io[1].gin + c);
            }
        }
        block_sum<2>(a, s_red);
        if (tid == 0) {
            double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
            if (C0.rd0) a00 = P.io[0].a0[row];
            if (C0.rd1) a01 = P.io[0].a1[row];
            if (C1.rd0) a10 = P.io[1].a0[row];
            if (C1.rd1) a11 = P.io[1].a1[row];
            if (PAIR) {
                const double2 old2 = P.self2[row];
                double2 nw = old2;
                if (act0) nw.x = row_epilogue(C0, a[0], old2.x, a00, a01, acc[0], acc[1]);
                if (act1) nw.y = row_epilogue(C1, a[1], old2.y, a10, a11, acc[2], acc[3]);
                P.self2[row] = nw;
            } else {
                if (act0) {
                    const double so = C0.rdself ? P.io[0].self[row] : 0.0;
                    P.io[0].self[row] = row_epilogue(C0, a[0], so, a00, a01, acc[0], acc[1]);
                }
                if (act1) {
                    const double so = C1.rdself ? P.io[1].self[row] : 0.0;
                    P.io[1].self[row] = row_epilogue(C1, a[1], so, a10, a11, acc[2], acc[3]);
                }
            }
            if (C0.wr0) P.io[0].a0[row] = a00;
            if (C0.wr1) P.io[0].a1[row] = a01;
            if (C1.wr0) P.io[1].a0[row] = a10;
            if (C1.wr1) P.io[1].a1[row] = a11;
        }
    }
    if (!use_state) return;

    // deterministic norms: one partial per CTA, the last CTA reduces them in a fixed order
    block_sum<4>(acc, s_red);
    if (tid == 0) {
        double *pp = P.partials + (size_t)cta * 4;
        pp[0] = acc[0]; pp[1] = acc[1]; pp[2] = acc[2]; pp[3] = acc[3];
        __threadfence();
        unsigned t = atomicAdd(P.counter, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double tot[4] = {0.0, 0.0, 0.0, 0.0};
    for (int i = tid; i < (int)gridDim.x; i += kStepThreads) {
        const double *pp = P.partials + (size_t)i * 4;
        tot[0] += __ldcg(pp + 0); tot[1] += __ldcg(pp + 1);
        tot[2] += __ldcg(pp + 2); tot[3] += __ldcg(pp + 3);
    }
    block_sum<4>(tot, s_red);
    if (tid == 0) {
        if (act0) finish_step(P.st[0], P.io[0].mode, tot[0], tot[1]);
        if (act1) finish_step(P.st[1], P.io[1].mode, tot[2], tot[3]);
        if (!P.st[0].active && !P.st[1].active) *P.done_flag = 1;
        *P.counter = 0;
        __threadfence();
    }
}

// ------------------------------------------------------------------------------------------------
// element-wise fused kernels (initialisation, MINRES / CGLS vector updates, output flush)
// ------------------------------------------------------------------------------------------------
struct EwParams {
    int op, slot, n, pair_slot;      // pair_slot: -1 plain, else column of the interleaved pair
    const double *in0;
    double *v0, *v1, *v2, *v3, *v4;
    double2 *pair;
    double c0;
    SlotState *st;
    double *partials;
    unsigned *counter;
    int *done_flag;
    int use_state;
};

__global__ void __launch_bounds__(kBlock) ew_kernel(EwParams P) {
    __shared__ double s_red[16];
    __shared__ int s_last;
    SlotState *S = P.st ? &P.st[P.slot] : nullptr;
    const bool is_init = (P.op == EW_INIT_LSQR || P.op == EW_INIT_CRAIG || P.op == EW_MINRES_INIT ||
                          P.op == EW_CGLS_INIT);
    const bool is_out = (P.op == EW_COPY || P.op == EW_CRAIG_FLUSH);
    if (P.use_state && !is_init && !is_out && !S->active) return;
    double acc[1] = {0.0};
    const int stride = gridDim.x * blockDim.x;
    // coefficients
    double alpha = 0, beta = 0, delta = 0, eps_ = 0, gamma = 1, phi = 0, xi = 0, c1 = 0,
           s1 = 0, sv = 0, lambda = 0, beta_c = 0;
    int iter = 0, pend = 0;
    if (S && !is_init) {
        alpha = S->alpha; beta = S->beta; delta = S->delta; eps_ = S->eps_;
        gamma = S->gamma; phi = S->phi; xi = S->xi; c1 = S->c1; s1 = S->s1; sv = S->sv;
        lambda = S->lambda; beta_c = S->beta_c; iter = S->iter; pend = S->pend;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
        switch (P.op) {
            case EW_COPY: P.v0[i] = P.c0 * P.in0[i]; break;
            case EW_INIT_LSQR: {
                double b = P.in0[i];
                double *dst = reinterpret_cast<double *>(P.pair + i) + P.pair_slot;
                *dst = b;
                acc[0] += b * b;
            } break;
            case EW_INIT_CRAIG: {
                double b = P.c0 * P.in0[i];
                double *dst = reinterpret_cast<double *>(P.pair + i) + P.pair_slot;
                *dst = b;
                P.v0[i] = 0.0;   // w
                P.v1[i] = 0.0;   // y
                acc[0] += b * b;
            } break;
            case EW_CRAIG_FLUSH: {
                // v0 = x, v1 = w2, pair = vhat ; out v2 = -(x + pending)
                double x = P.v0[i];
                if (pend) {
                    double vp = (reinterpret_cast<const double *>(P.pair + i))[P.pair_slot] * sv;
                    if (lambda > 0) { x = x + (xi * c1) * vp; x = x + (xi * s1) * P.v1[i]; }
                    else x += xi * vp;
                }
                P.v2[i] = -x;
            } break;
            case EW_MINRES_INIT: {
                double b = P.in0[i];
                P.v0[i] = b;     // r1
                P.v1[i] = b;     // r2
                P.v2[i] = 0.0;   // w1
                P.v3[i] = 0.0;   // w2
                P.v4[i] = 0.0;   // x
                acc[0] += b * b;
            } break;
            case EW_MINRES_E1: {
                // in0 = y(pre) ; v0 = r1, v1 = r2, v2 = w1-role, v3 = w2-role ; iter already incremented
                double r2 = P.v1[i];
                double y = P.in0[i] - (alpha / beta) * r2;
                double w;
                if (iter == 1) {
                    w = P.v3[i] + (1.0 / beta) * r2;
                    P.v3[i] = w;
                } else {
                    double w1 = P.v2[i];
                    if (iter >= 3) w1 *= -eps_;
                    w1 -= delta * P.v3[i];
                    w = w1 + (1.0 / beta) * r2;
                    P.v2[i] = w;
                }
                P.v0[i] = r2;
                P.v1[i] = y;
                acc[0] += y * y;
            } break;
            case EW_MINRES_E2: {
                // v2 = the w just written (role resolved by the host), v4 = x
                double w = P.v2[i] * (1.0 / gamma);
                P.v2[i] = w;
                double x = P.v4[i] + phi * w;
                P.v4[i] = x;
                acc[0] += x * x;
            } break;
            case EW_CGLS_INIT: {
                double b = P.in0[i];
                P.v0[i] = b;     // r
                acc[0] += b * b;
            } break;
            case EW_CGLS_EN: {   // r -= alpha q
                double r = P.v0[i] - alpha * P.v1[i];
                P.v0[i] = r;
                acc[0] += r * r;
            } break;
            case EW_CGLS_EM: {   // p = s + beta p
                double p = P.v0[i] + beta_c * P.v1[i];
                P.v1[i] = p;
                acc[0] += p * p;
            } break;
            default: break;
        }
    }
    if (!P.use_state || is_out) return;
    block_sum<1>(acc, s_red);
    if (threadIdx.x == 0) {
        P.partials[blockIdx.x] = acc[0];
        __threadfence();
        unsigned t = atomicAdd(P.counter, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    double tot[1] = {0.0};
    for (int i = threadIdx.x; i < (int)gridDim.x; i += kBlock) tot[0] += __ldcg(P.partials + i);
    block_sum<1>(tot, s_red);
    if (threadIdx.x == 0) {
        finish_ew(*S, P.op, tot[0]);
        if (!P.st[0].active && !P.st[1].active) *P.done_flag = 1;
        *P.counter = 0;
        __threadfence();
    }
}

__global__ void gather_vals_kernel(int nnz, const int *perm, const double *coo, double *vx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) { const int p = perm[i]; vx[i] = (p >= 0) ? coo[p] : 0.0; }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct IterWs {
    DevBuf<double2> Gnm;                 // one allocation: [Gn | Gm] so a single L2 persisting window covers both
    struct View { double2 *p = nullptr; size_t n = 0;
                  void zero(cudaStream_t s) { if (p) FPSB_CUDA(cudaMemsetAsync(p, 0, n * sizeof(double2), s)); } };
    View Gn, Gm;                         // interleaved Golub-Kahan pairs (n-space, m-space)
    DevBuf<double> an[2][2];             // n-space aux per slot (CRAIG x, w2 ; CGLS r, q ; MINRES t)
    DevBuf<double> am[2][5];             // m-space aux per slot
    DevBuf<double> ym;                   // MINRES y / CGLS s
    DevBuf<SlotState> st;
    DevBuf<double> partials;
    DevBuf<unsigned> counter;
    DevBuf<int> done;
    int *h_done = nullptr;               // pinned
    SlotState *h_st = nullptr;           // pinned [2]
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaEvent_t pev[2] = {nullptr, nullptr};   // profile: Krylov loop region
    int64_t prof_launch0 = 0, prof_launch1 = 0;
    bool prof_armed = false;
    int ew_grid = 0;
};

static void build_csr_host(int nrows, int ncols, int64_t nnz, const int64_t *ri, const int64_t *cj,
                           std::vector<int> &rp, std::vector<int> &ci, std::vector<int> &perm) {
    rp.assign((size_t)nrows + 1, 0);
    for (int64_t k = 0; k < nnz; ++k) rp[(size_t)ri[k] + 1]++;
    for (int i = 0; i < nrows; ++i) rp[i + 1] += rp[i];
    ci.resize((size_t)nnz);
    perm.resize((size_t)nnz);
    // stable counting sort by column first, then by row => rows hold ascending columns
    std::vector<int> cp((size_t)ncols + 1, 0);
    for (int64_t k = 0; k < nnz; ++k) cp[(size_t)cj[k] + 1]++;
    for (int j = 0; j < ncols; ++j) cp[j + 1] += cp[j];
    std::vector<int> bycol((size_t)nnz);
    for (int64_t k = 0; k < nnz; ++k) bycol[(size_t)cp[(size_t)cj[k]]++] = (int)k;
    std::vector<int> pos(rp.begin(), rp.end() - 1);
    for (int64_t t = 0; t < nnz; ++t) {
        int k = bycol[(size_t)t];
        int q = pos[(size_t)ri[k]]++;
        ci[(size_t)q] = (int)cj[k];
        perm[(size_t)q] = k;
    }
}

constexpr int kLongRow = 1024;   // rows longer than this leave the SELL part
constexpr int kSigma = 2048;     // sorting window (rows) of SELL-32-sigma

// CSR -> SELL-32-sigma (+ CSR of the long rows) and upload
static void upload_sell(Handle *h, CsrDev &M, int nrows, int ncols, const std::vector<int> &rp,
                        const std::vector<int> &ci, const std::vector<int> &perm) {
    M.nrows = nrows; M.ncols = ncols; M.nnz = (int64_t)ci.size();
    std::vector<int> rowidx, sl_off(1, 0), long_row, long_rp(1, 0), long_col, long_perm;
    std::vector<int> order;
    for (int w0 = 0; w0 < nrows; w0 += kSigma) {
        const int w1 = std::min(nrows, w0 + kSigma);
        order.clear();
        for (int r = w0; r < w1; ++r) {
            const int len = rp[(size_t)r + 1] - rp[(size_t)r];
            if (len > kLongRow) {
                long_row.push_back(r);
                for (int p = rp[(size_t)r]; p < rp[(size_t)r + 1]; ++p) { long_col.push_back(ci[(size_t)p]); long_perm.push_back(perm[(size_t)p]); }
                long_rp.push_back((int)long_col.size());
            } else order.push_back(r);
        }
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            return (rp[(size_t)a + 1] - rp[(size_t)a]) > (rp[(size_t)b + 1] - rp[(size_t)b]);
        });
        for (size_t i = 0; i < order.size(); i += 32) {
            int width = 0;
            for (size_t l = 0; l < 32; ++l) {
                const int r = (i + l < order.size()) ? order[i + l] : -1;
                rowidx.push_back(r);
                if (r >= 0) width = std::max(width, rp[(size_t)r + 1] - rp[(size_t)r]);
            }
            sl_off.push_back(sl_off.back() + width * 32);
        }
    }
    const int nslice = (int)sl_off.size() - 1;
    const size_t padded = (size_t)sl_off.back();
    std::vector<int> scol(padded, 0), sperm(padded, -1);
    for (int sidx = 0; sidx < nslice; ++sidx) {
        const int off = sl_off[(size_t)sidx], width = (sl_off[(size_t)sidx + 1] - off) / 32;
        for (int l = 0; l < 32; ++l) {
            const int r = rowidx[(size_t)sidx * 32 + l];
            int len = 0, base = 0;
            if (r >= 0) { base = rp[(size_t)r]; len = rp[(size_t)r + 1] - base; }
            int lastc = (len > 0) ? ci[(size_t)(base + len - 1)] : 0;
            for (int j = 0; j < width; ++j) {
                const size_t q = (size_t)off + (size_t)j * 32 + l;
                if (j < len) { scol[q] = ci[(size_t)(base + j)]; sperm[q] = perm[(size_t)(base + j)]; }
                else { scol[q] = lastc; sperm[q] = -1; }     // padding: value 0, local column
            }
        }
    }
    M.nslice = nslice;
    M.padded = (int64_t)padded;
    M.nlong = (int)long_row.size();
    M.sl_off.from(sl_off, h->stream);
    M.rowidx.from(rowidx, h->stream);
    M.scol.from(scol, h->stream);
    M.sperm.from(sperm, h->stream);
    M.sval.alloc(padded + 8);
    M.sval.zero(h->stream);
    M.long_row.from(long_row, h->stream);
    M.long_rp.from(long_rp, h->stream);
    M.long_col.from(long_col, h->stream);
    M.long_perm.from(long_perm, h->stream);
    M.long_val.alloc(long_col.size() + 8);
    M.long_val.zero(h->stream);
    M.grid_sell = nslice > 0 ? std::max(1, std::min((nslice + 7) / 8, 2 * h->num_sms)) : 0;
    {
        // slice metadata of a CTA must fit its shared-memory staging area; keep whole waves of
        // resident CTAs (2 per SM) so there is no ragged tail wave
        const int resident = 2 * h->num_sms;
        const int need = (nslice + kMaxSlicesPerCta - 1) / kMaxSlicesPerCta;
        if (need > M.grid_sell) M.grid_sell = ((need + resident - 1) / resident) * resident;
    }
    M.grid = M.grid_sell + M.nlong;
    {
        // contiguous chunks of slices per warp, balanced by streamed rows (+1 per slice for the
        // epilogue); wchunk[w] = first slice of warp w
        const int nw = M.grid_sell * (kBlock / 32);
        std::vector<int> wchunk((size_t)nw + 1, nslice);
        std::vector<int64_t> cost((size_t)nslice + 1, 0);
        for (int i = 0; i < nslice; ++i) cost[(size_t)i + 1] = cost[(size_t)i] + (sl_off[(size_t)i + 1] - sl_off[(size_t)i]) / 32 + 2;
        int cur = 0;
        for (int w = 0; w < nw; ++w) {
            wchunk[(size_t)w] = cur;
            const int64_t target = nw > 0 ? cost[(size_t)nslice] * (w + 1) / nw : 0;
            while (cur < nslice && cost[(size_t)cur + 1] <= target) cur++;
        }
        wchunk[(size_t)nw] = nslice;
        if (nw > 0) wchunk[0] = 0;
        M.wchunk.from(wchunk, h->stream);
    }
}

void csr_build(Handle *h) {
    const int m = (int)h->ncon, n = (int)h->nvar;
    {
        int dev = 0, sms = 148;
        FPSB_CUDA(cudaGetDevice(&dev));
        FPSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        h->num_sms = sms;
        FPSB_CUDA(cudaFuncSetAttribute(gk_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingBytes));
        FPSB_CUDA(cudaFuncSetAttribute(gk_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kRingBytes));
    }
    std::vector<int> rp, ci, perm;
    build_csr_host(m, n, h->nnzj, h->jrow.data(), h->jcol.data(), rp, ci, perm);
    upload_sell(h, h->A, m, n, rp, ci, perm);
    build_csr_host(n, m, h->nnzj, h->jcol.data(), h->jrow.data(), rp, ci, perm);
    upload_sell(h, h->At, n, m, rp, ci, perm);
    h->coo_vals.alloc((size_t)h->nnzj + 8);
    FPSB_CUDA(cudaStreamSynchronize(h->stream));
}

void csr_refresh_values(Handle *h) {
    for (CsrDev *M : {&h->A, &h->At}) {
        if (M->padded > 0) {
            int grid = (int)((M->padded + 255) / 256);
            gather_vals_kernel<<<grid, 256, 0, h->stream>>>((int)M->padded, M->sperm.p, h->coo_vals.p, M->sval.p);
            h->launches += 1;
        }
        const int nl = (int)(M->long_col.n > 8 ? M->long_col.n - 8 : 0);
        if (M->nlong > 0 && nl > 0) {
            gather_vals_kernel<<<(nl + 255) / 256, 256, 0, h->stream>>>(nl, M->long_perm.p, h->coo_vals.p, M->long_val.p);
            h->launches += 1;
        }
    }
    FPSB_CUDA(cudaGetLastError());
}

static void fill_csr(StepParams &P, const CsrDev &M) {
    P.wchunk = M.wchunk.p;
    P.sl_off = M.sl_off.p; P.rowidx = M.rowidx.p; P.scol = M.scol.p; P.sval = M.sval.p;
    P.nslice = M.nslice; P.grid_sell = M.grid_sell;
    P.long_row = M.long_row.p; P.long_rp = M.long_rp.p; P.long_col = M.long_col.p; P.long_val = M.long_val.p;
    P.nrows = M.nrows;
}

// y = Op x for 1 or 2 plain columns (columns contiguous in memory)
void spmv_plain(Handle *h, bool transpose, const double *x, double *y, int ncols_rhs) {
    const CsrDev &M = transpose ? h->At : h->A;
    if (M.nrows == 0 || M.grid == 0) return;
    StepParams P{};
    fill_csr(P, M);
    for (int s = 0; s < 2; ++s) {
        P.io[s].mode = (s < ncols_rhs) ? MD_PLAIN : MD_NONE;
        P.io[s].gin = x + (size_t)s * M.ncols;
        P.io[s].self = y + (size_t)s * M.nrows;
        P.io[s].a0 = nullptr;
        P.io[s].c0 = 1.0; P.io[s].c1 = 0.0;
    }
    gk_step_kernel<false><<<M.grid, kStepThreads, kRingBytes, h->stream>>>(P, 0);
    h->launches += 1;
    FPSB_CUDA(cudaGetLastError());
}

void iter_setup(Handle *h) {
    if (h->iter) return;
    IterWs *W = new IterWs();
    h->iter = W;
    const size_t n = (size_t)h->nvar, m = (size_t)h->ncon;
    {
        const size_t gn = (n + 4 + 7) & ~(size_t)7, gm = (m + 4 + 7) & ~(size_t)7;
        W->Gnm.alloc(gn + gm);
        W->Gn.p = W->Gnm.p; W->Gn.n = gn;
        W->Gm.p = W->Gnm.p + gn; W->Gm.n = gm;
        // keep the gathered vectors resident in L2 across the 240 MB matrix stream of every iteration
        int dev = 0, maxp = 0, maxw = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, dev);
        cudaDeviceGetAttribute(&maxw, cudaDevAttrMaxAccessPolicyWindowSize, dev);
        const size_t bytes = (gn + gm) * sizeof(double2);
        if (maxp > 0 && maxw > 0) {
            const size_t want = std::min(bytes, (size_t)maxp);
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof(attr));
            attr.accessPolicyWindow.base_ptr = (void *)W->Gnm.p;
            attr.accessPolicyWindow.num_bytes = std::min(bytes, (size_t)maxw);
            attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)std::max<size_t>(bytes, 1));
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (cudaStreamSetAttribute(h->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
        }
        cudaGetLastError();
    }
    for (int s = 0; s < 2; ++s) {
        for (int k = 0; k < 2; ++k) W->an[s][k].alloc(n + 4);
        for (int k = 0; k < 5; ++k) W->am[s][k].alloc(m + 4);
    }
    W->ym.alloc(m + 4);
    W->st.alloc(2);
    int maxblk = std::max(h->A.grid, h->At.grid);
    W->ew_grid = 148 * 4;
    W->partials.alloc((size_t)std::max(maxblk, W->ew_grid) * 4 + 16);
    W->counter.alloc(4);
    W->done.alloc(4);
    W->counter.zero(h->stream);
    W->done.zero(h->stream);
    FPSB_CUDA(cudaMallocHost((void **)&W->h_done, 2 * sizeof(int)));
    FPSB_CUDA(cudaMallocHost((void **)&W->h_st, 2 * sizeof(SlotState)));
    FPSB_CUDA(cudaEventCreateWithFlags(&W->ev[0], cudaEventDisableTiming));
    FPSB_CUDA(cudaEventCreateWithFlags(&W->ev[1], cudaEventDisableTiming));
    FPSB_CUDA(cudaEventCreate(&W->pev[0]));
    FPSB_CUDA(cudaEventCreate(&W->pev[1]));
    FPSB_CUDA(cudaStreamSynchronize(h->stream));
}

void iter_free(Handle *h) {
    if (!h->iter) return;
    IterWs *W = h->iter;
    if (W->h_done) cudaFreeHost(W->h_done);
    if (W->h_st) cudaFreeHost(W->h_st);
    if (W->ev[0]) cudaEventDestroy(W->ev[0]);
    if (W->ev[1]) cudaEventDestroy(W->ev[1]);
    if (W->pev[0]) cudaEventDestroy(W->pev[0]);
    if (W->pev[1]) cudaEventDestroy(W->pev[1]);
    delete W;
    h->iter = nullptr;
}

// ---- slot configuration (host fills a SlotState, uploaded before the solve) ---------------------
static SlotState make_lsqr(double lambda, double atol, double rtol, int64_t itmax, int64_t m_op, int64_t n_op) {
    SlotState S;
    memset(&S, 0, sizeof(S));
    S.algo = ALGO_LSQR;
    S.lambda = lambda; S.atol = atol; S.rtol = rtol;
    S.axtol = kSqrtEps; S.btol = kSqrtEps; S.etol = kSqrtEps; S.ctol = kSqrtEps;   // conlim = 1/sqrt(eps)
    int64_t im = itmax == 0 ? m_op + n_op : itmax;
    S.itmax = (int)std::min<int64_t>(im, 2000000000);
    S.active = 1;
    S.mscale = 1.0;
    return S;
}
static SlotState make_craig(double delta, double atol, double rtol, double btol, double conlim,
                            int64_t itmax, int64_t m_op, int64_t n_op) {
    SlotState S;
    memset(&S, 0, sizeof(S));
    S.algo = ALGO_CRAIG;
    S.sqd = delta != 0.0;
    S.lambda = S.sqd ? 1.0 : 0.0;
    S.mscale = S.sqd ? 1.0 / delta : 1.0;
    S.atol = atol; S.rtol = rtol; S.btol = btol;
    S.ctol = conlim > 0 ? 1.0 / conlim : 0.0;
    int64_t im = itmax == 0 ? m_op + n_op : itmax;
    S.itmax = (int)std::min<int64_t>(im, 2000000000);
    S.active = 1;
    return S;
}
static SlotState make_minres(double lambda, double atol, double rtol, double etol, double conlim,
                             int64_t itmax, int64_t n_op) {
    SlotState S;
    memset(&S, 0, sizeof(S));
    S.algo = ALGO_MINRES;
    S.lambda = lambda; S.atol = atol; S.rtol = rtol; S.etol = etol;
    S.ctol = conlim > 0 ? 1.0 / conlim : 0.0;
    int64_t im = itmax == 0 ? 2 * n_op : itmax;
    S.itmax = (int)std::min<int64_t>(im, 2000000000);
    S.active = 1;
    S.mscale = 1.0;
    return S;
}
static SlotState make_cgls(double lambda, double atol, double rtol, int64_t itmax, int64_t m_op, int64_t n_op) {
    SlotState S;
    memset(&S, 0, sizeof(S));
    S.algo = ALGO_CGLS;
    S.lambda = lambda; S.atol = atol; S.rtol = rtol;
    int64_t im = itmax == 0 ? m_op + n_op : itmax;
    S.itmax = (int)std::min<int64_t>(im, 2000000000);
    S.active = 1;
    S.mscale = 1.0;
    return S;
}
static SlotState make_none() {
    SlotState S;
    memset(&S, 0, sizeof(S));
    S.algo = ALGO_NONE;
    S.active = 0;
    S.solved = 1;
    return S;
}

static void stats_from(const SlotState &S, fpsb_krylov_stats &o) {
    o.niter = S.iter; o.solved = S.solved; o.inconsistent = S.inconsistent; o.status = S.status;
    o.pad_ = 0;
    o.rnorm = S.rNorm; o.arnorm = S.ArNorm; o.anorm = S.Anorm; o.acond = S.Acond; o.xnorm = S.xNorm;
}

struct Engine {
    Handle *h;
    IterWs *W;
    StepParams base_m, base_n;   // M: rows of A (m-space rows), N: rows of A' (n-space rows)
    Engine(Handle *hh) : h(hh), W(hh->iter) {
        memset(&base_m, 0, sizeof(base_m));
        memset(&base_n, 0, sizeof(base_n));
        fill_csr(base_m, h->A);
        fill_csr(base_n, h->At);
        for (StepParams *P : {&base_m, &base_n}) {
            P->st = W->st.p; P->partials = W->partials.p; P->counter = W->counter.p; P->done_flag = W->done.p;
        }
        base_m.gin2 = W->Gn.p; base_m.self2 = W->Gm.p;
        base_n.gin2 = W->Gm.p; base_n.self2 = W->Gn.p;
    }
    void begin(const SlotState &s0, const SlotState &s1) {
        SlotState hs[2] = {s0, s1};
        memcpy(W->h_st, hs, sizeof(hs));
        FPSB_CUDA(cudaMemcpyAsync(W->st.p, W->h_st, sizeof(hs), cudaMemcpyHostToDevice, h->stream));
        FPSB_CUDA(cudaMemsetAsync(W->done.p, 0, sizeof(int), h->stream));
        FPSB_CUDA(cudaMemsetAsync(W->counter.p, 0, sizeof(unsigned), h->stream));
    }
    void step(bool mspace, bool pair, const SlotIO &io0, const SlotIO &io1) {
        StepParams P = mspace ? base_m : base_n;
        P.io[0] = io0; P.io[1] = io1;
        const CsrDev &M = mspace ? h->A : h->At;
        if (M.grid == 0) return;
        if (pair) gk_step_kernel<true><<<M.grid, kStepThreads, kRingBytes, h->stream>>>(P, 1);
        else gk_step_kernel<false><<<M.grid, kStepThreads, kRingBytes, h->stream>>>(P, 1);
        h->launches += 1;
    }
    void ew(int op, int slot, int n, const double *in0, double *v0, double *v1, double *v2, double *v3,
            double *v4, double2 *pair, int pair_slot, double c0, int use_state = 1) {
        if (n == 0 && !(use_state)) return;
        EwParams P{};
        P.op = op; P.slot = slot; P.n = n; P.pair_slot = pair_slot; P.in0 = in0;
        P.v0 = v0; P.v1 = v1; P.v2 = v2; P.v3 = v3; P.v4 = v4; P.pair = pair; P.c0 = c0;
        P.st = W->st.p; P.partials = W->partials.p; P.counter = W->counter.p; P.done_flag = W->done.p;
        P.use_state = use_state;
        int grid = std::max(1, std::min(W->ew_grid, (n + kBlock - 1) / kBlock));
        ew_kernel<<<grid, kBlock, 0, h->stream>>>(P);
        h->launches += 1;
    }
    // run `body(k)` (k = 1, 2, ...) until the device reports every slot stopped
    template <class F>
    void loop(F body, int chunk) {
        int k = 0, pending = 0, slot = 0;
        int64_t hard_cap = (int64_t)4000000000LL;
        bool done = false;
        while (!done && k < hard_cap) {
            for (int c = 0; c < chunk; ++c) body(++k);
            FPSB_CUDA(cudaMemcpyAsync(&W->h_done[slot], W->done.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            FPSB_CUDA(cudaEventRecord(W->ev[slot], h->stream));
            pending++;
            if (pending == 2) {
                int prev = slot ^ 1;
                FPSB_CUDA(cudaEventSynchronize(W->ev[prev]));
                if (W->h_done[prev] != 0) done = true;
                pending--;
            }
            slot ^= 1;
        }
        FPSB_CUDA(cudaStreamSynchronize(h->stream));
        FPSB_CUDA(cudaGetLastError());
        if (W->h_done[0] < 0 || W->h_done[1] < 0) {
            set_error("TMA/mbarrier wait timed out inside gk_step_kernel");
            throw CudaFail{FPSB_ECUDA};
        }
    }
    // CUDA-event bracket around the Krylov loop (first step kernel .. last chunk), for the roofline
    void mark_begin() {
        FPSB_CUDA(cudaEventRecord(W->pev[0], h->stream));
        W->prof_launch0 = h->launches;
        W->prof_armed = true;
    }
    void mark_end() {
        FPSB_CUDA(cudaEventRecord(W->pev[1], h->stream));
        W->prof_launch1 = h->launches;
    }
    void fetch(fpsb_krylov_stats *st) {
        FPSB_CUDA(cudaMemcpyAsync(W->h_st, W->st.p, 2 * sizeof(SlotState), cudaMemcpyDeviceToHost, h->stream));
        FPSB_CUDA(cudaStreamSynchronize(h->stream));
        stats_from(W->h_st[0], st[0]);
        stats_from(W->h_st[1], st[1]);
        if (W->prof_armed) {
            float ms = 0.f;
            FPSB_CUDA(cudaEventElapsedTime(&ms, W->pev[0], W->pev[1]));
            h->prof_loop_ms = ms;
            h->prof_step_launches = W->prof_launch1 - W->prof_launch0;
            W->prof_armed = false;
        }
    }
};

static SlotIO io_none() { SlotIO io{}; io.mode = MD_NONE; return io; }
static SlotIO io_mode(int mode, double *a0 = nullptr, double *a1 = nullptr, double *a2 = nullptr) {
    SlotIO io{};
    io.mode = mode; io.a0 = a0; io.a1 = a1; io.a2 = a2; io.c0 = 1.0; io.c1 = 0.0;
    return io;
}

// p = rhs - A' q  (n-space), plain kernel
static void residual_p(Engine &E, const double *rhs, const double *q, double *p) {
    Handle *h = E.h;
    if (h->At.grid == 0) return;
    StepParams P = E.base_n;
    P.io[0] = io_mode(MD_PLAIN, const_cast<double *>(rhs));
    P.io[0].gin = q; P.io[0].self = p; P.io[0].c0 = -1.0; P.io[0].c1 = 1.0;
    P.io[1] = io_none();
    gk_step_kernel<false><<<h->At.grid, kStepThreads, kRingBytes, h->stream>>>(P, 0);
    h->launches += 1;
}

static const int kChunk = 4;

// LSQR on A' occupies: Gn[:,slot] = u, Gm[:,slot] = v, am[slot][0] = w, am[slot][1] = x
static void lsqr_init(Engine &E, int slot, const double *rhs) {
    E.ew(EW_INIT_LSQR, slot, (int)E.h->nvar, rhs, nullptr, nullptr, nullptr, nullptr, nullptr, E.W->Gn.p, slot, 1.0);
}

void iter_solve_two_mixed(Handle *h, double delta, const double *rhs1, const double *rhs2, double *p1,
                          double *q1, double *p2, double *q2, fpsb_krylov_stats *st) {
    iter_setup(h);
    Engine E(h);
    IterWs *W = h->iter;
    const fpsb_iter_opts &o = h->iopts;
    const int64_t n = h->nvar, m = h->ncon;
    E.begin(make_lsqr(sqrt(delta), o.ls_atol, o.ls_rtol, o.ls_itmax, n, m),
            make_craig(delta, o.ln_atol, o.ln_rtol, o.ln_btol, o.ln_conlim, o.ln_itmax, m, n));
    W->Gn.zero(h->stream); W->Gm.zero(h->stream);
    W->an[1][0].zero(h->stream); W->an[1][1].zero(h->stream);   // CRAIG x, w2
    W->am[0][1].zero(h->stream);                                // LSQR x
    lsqr_init(E, 0, rhs1);
    E.ew(EW_INIT_CRAIG, 1, (int)m, rhs2, W->am[1][0].p, W->am[1][1].p, nullptr, nullptr, nullptr, W->Gm.p, 1, -1.0);
    SlotIO l_init = io_mode(MD_LSQR_INIT_M);
    SlotIO l_u = io_mode(MD_LSQR_U);
    SlotIO l_v = io_mode(MD_LSQR_V, W->am[0][0].p, W->am[0][1].p);
    SlotIO c_v = io_mode(MD_CRAIG_V, W->an[1][0].p, W->an[1][1].p);
    SlotIO c_u = io_mode(MD_CRAIG_U, W->am[1][0].p, W->am[1][1].p);
    E.mark_begin();
    E.step(true, true, l_init, io_none());
    E.loop([&](int) {
        E.step(false, true, l_u, c_v);
        E.step(true, true, l_v, c_u);
    }, kChunk);
    E.mark_end();
    // outputs
    E.ew(EW_COPY, 0, (int)m, W->am[0][1].p, q1, nullptr, nullptr, nullptr, nullptr, nullptr, -1, 1.0, 0);
    residual_p(E, rhs1, q1, p1);
    E.ew(EW_CRAIG_FLUSH, 1, (int)n, nullptr, W->an[1][0].p, W->an[1][1].p, p2, nullptr, nullptr, W->Gn.p, 1, 1.0, 1);
    E.ew(EW_COPY, 1, (int)m, W->am[1][1].p, q2, nullptr, nullptr, nullptr, nullptr, nullptr, -1, 1.0, 0);
    E.fetch(st);
}

void iter_solve_two_least_squares(Handle *h, double delta, const double *rhs1, const double *rhs2,
                                  double *p1, double *q1, double *p2, double *q2,
                                  fpsb_krylov_stats *st) {
    iter_setup(h);
    Engine E(h);
    IterWs *W = h->iter;
    const fpsb_iter_opts &o = h->iopts;
    const int64_t n = h->nvar, m = h->ncon;
    SlotState s = make_lsqr(sqrt(delta), o.ls_atol, o.ls_rtol, o.ls_itmax, n, m);
    E.begin(s, s);
    W->Gn.zero(h->stream); W->Gm.zero(h->stream);
    W->am[0][1].zero(h->stream); W->am[1][1].zero(h->stream);
    lsqr_init(E, 0, rhs1);
    lsqr_init(E, 1, rhs2);
    SlotIO init0 = io_mode(MD_LSQR_INIT_M), init1 = io_mode(MD_LSQR_INIT_M);
    SlotIO u0 = io_mode(MD_LSQR_U), u1 = io_mode(MD_LSQR_U);
    SlotIO v0 = io_mode(MD_LSQR_V, W->am[0][0].p, W->am[0][1].p);
    SlotIO v1 = io_mode(MD_LSQR_V, W->am[1][0].p, W->am[1][1].p);
    E.mark_begin();
    E.step(true, true, init0, init1);
    E.loop([&](int) {
        E.step(false, true, u0, u1);
        E.step(true, true, v0, v1);
    }, kChunk);
    E.mark_end();
    E.ew(EW_COPY, 0, (int)m, W->am[0][1].p, q1, nullptr, nullptr, nullptr, nullptr, nullptr, -1, 1.0, 0);
    E.ew(EW_COPY, 1, (int)m, W->am[1][1].p, q2, nullptr, nullptr, nullptr, nullptr, nullptr, -1, 1.0, 0);
    // p_i = rhs_i - A' q_i : one two-column SpMM
    if (h->At.grid) {
        StepParams P = E.base_n;
        P.io[0] = io_mode(MD_PLAIN, const_cast<double *>(rhs1));
        P.io[0].gin = q1; P.io[0].self = p1; P.io[0].c0 = -1.0; P.io[0].c1 = 1.0;
        P.io[1] = io_mode(MD_PLAIN, const_cast<double *>(rhs2));
        P.io[1].gin = q2; P.io[1].self = p2; P.io[1].c0 = -1.0; P.io[1].c1 = 1.0;
        gk_step_kernel<false><<<h->At.grid, kStepThreads, kRingBytes, h->stream>>>(P, 0);
        h->launches += 1;
    }
    E.fetch(st);
}

// MINRES on (A A' + lambda I) in slot `slot` (non-pair kernels), result copied to `out`
static void run_minres(Engine &E, int slot, const double *rhs, double *out) {
    Handle *h = E.h;
    IterWs *W = E.W;
    const int m = (int)h->ncon;
    double *r1 = W->am[slot][0].p, *r2 = W->am[slot][1].p, *wa = W->am[slot][2].p, *wb = W->am[slot][3].p,
           *x = W->am[slot][4].p, *y = W->ym.p, *t = W->an[slot][0].p;
    E.ew(EW_MINRES_INIT, slot, m, rhs, r1, r2, wa, wb, x, nullptr, -1, 1.0);
    double *w1 = wa, *w2 = wb;
    SlotIO none = io_none();
    E.loop([&](int k) {
        // t = A' r2 ; y = (A t + lambda r2)/beta - (beta/oldbeta) r1 ; alpha = r2'y / beta
        SlotIO n_io = io_mode(MD_PLAIN);
        n_io.gin = r2; n_io.self = t; n_io.c0 = 1.0;
        SlotIO m_io = io_mode(MD_MINRES_M, r2, r1);
        m_io.gin = t; m_io.self = y;
        if (slot == 0) { E.step(false, false, n_io, none); E.step(true, false, m_io, none); }
        else { E.step(false, false, none, n_io); E.step(true, false, none, m_io); }
        E.ew(EW_MINRES_E1, slot, m, y, r1, r2, w1, w2, nullptr, nullptr, -1, 1.0);
        double *wcur = (k == 1) ? w2 : w1;
        E.ew(EW_MINRES_E2, slot, m, nullptr, nullptr, nullptr, wcur, nullptr, x, nullptr, -1, 1.0);
        if (k >= 2) std::swap(w1, w2);
    }, 2);
    E.ew(EW_COPY, slot, m, x, out, nullptr, nullptr, nullptr, nullptr, nullptr, -1, 1.0, 0);
}

// CGLS on A' (min ||b - A' x||^2 + lambda ||x||^2) in slot `slot`, result copied to `out`
static void run_cgls(Engine &E, int slot, const double *rhs, double *out) {
    Handle *h = E.h;
    IterWs *W = E.W;
    const int m = (int)h->ncon, n = (int)h->nvar;
    double *r = W->an[slot][0].p, *q = W->an[slot][1].p;
    double *x = W->am[slot][0].p, *p = W->am[slot][1].p, *s = W->am[slot][2].p;
    W->am[slot][0].zero(h->stream);
    E.ew(EW_CGLS_INIT, slot, n, rhs, r, nullptr, nullptr, nullptr, nullptr, nullptr, -1, 1.0);
    SlotIO none = io_none();
    {
        SlotIO io = io_mode(MD_CGLS_INIT_M);
        io.gin = r; io.self = p;
        if (slot == 0) E.step(true, false, io, none); else E.step(true, false, none, io);
    }
    E.loop([&](int) {
        SlotIO n_io = io_mode(MD_CGLS_N);
        n_io.gin = p; n_io.self = q;
        if (slot == 0) E.step(false, false, n_io, none); else E.step(false, false, none, n_io);
        E.ew(EW_CGLS_EN, slot, n, nullptr, r, q, nullptr, nullptr, nullptr, nullptr, -1, 1.0);
        SlotIO m_io = io_mode(MD_CGLS_M, x, p);
        m_io.gin = r; m_io.self = s;
        if (slot == 0) E.step(true, false, m_io, none); else E.step(true, false, none, m_io);
        E.ew(EW_CGLS_EM, slot, m, nullptr, s, p, nullptr, nullptr, nullptr, nullptr, -1, 1.0);
    }, 2);
    E.ew(EW_COPY, slot, m, x, out, nullptr, nullptr, nullptr, nullptr, nullptr, -1, 1.0, 0);
}

void iter_solve_two_extras(Handle *h, double delta, const double *rhs1, const double *rhs2, double *u1,
                           double *u2, fpsb_krylov_stats *st, bool ldlt_variant) {
    iter_setup(h);
    Engine E(h);
    IterWs *W = h->iter;
    const fpsb_iter_opts &o = h->iopts;
    const int64_t n = h->nvar, m = h->ncon;
    const double tau = std::max(delta, 1e-14);
    fpsb_krylov_stats tmp[2];
    if (!ldlt_variant) {
        // LSQR(A', rhs1, lambda = sqrt(tau))   src/solve_linear_system.jl:53
        E.begin(make_lsqr(sqrt(tau), o.ls_atol, o.ls_rtol, o.ls_itmax, n, m), make_none());
        W->Gn.zero(h->stream); W->Gm.zero(h->stream);
        W->am[0][1].zero(h->stream);
        lsqr_init(E, 0, rhs1);
        SlotIO l_init = io_mode(MD_LSQR_INIT_M), l_u = io_mode(MD_LSQR_U);
        SlotIO l_v = io_mode(MD_LSQR_V, W->am[0][0].p, W->am[0][1].p);
        E.step(true, true, l_init, io_none());
        E.loop([&](int) {
            E.step(false, true, l_u, io_none());
            E.step(true, true, l_v, io_none());
        }, kChunk);
        E.ew(EW_COPY, 0, (int)m, W->am[0][1].p, u1, nullptr, nullptr, nullptr, nullptr, nullptr, -1, 1.0, 0);
        E.fetch(tmp);
        st[0] = tmp[0];
        // MINRES(A A', rhs2, lambda = tau, ne_*)   src/solve_linear_system.jl:60-70
        E.begin(make_none(), make_minres(tau, o.ne_atol, o.ne_rtol, o.ne_etol, o.ne_conlim, o.ne_itmax, m));
        run_minres(E, 1, rhs2, u2);
        E.fetch(tmp);
        st[1] = tmp[1];
    } else {
        // cgls(A', rhs1, lambda = tau) ; minres(A A', rhs2, lambda = tau) — Krylov.jl defaults
        E.begin(make_cgls(tau, kSqrtEps, kSqrtEps, 0, n, m), make_none());
        run_cgls(E, 0, rhs1, u1);
        E.fetch(tmp);
        st[0] = tmp[0];
        E.begin(make_none(), make_minres(tau, kSqrtEps / 100, kSqrtEps / 100, kSqrtEps, 1.0 / kSqrtEps, 0, m));
        run_minres(E, 1, rhs2, u2);
        E.fetch(tmp);
        st[1] = tmp[1];
    }
}

}  // namespace fpsb
