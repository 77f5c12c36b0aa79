/* Looped C restatements of the other BASELINE.json configurations at benchmark size — TEST/BASELINE
 * INFRASTRUCTURE ONLY (the product never links or loads this file).
 *
 * Each function restates, per grid point, what the reference's discretisation produces for that
 * problem (the Python oracle oracle/discretize.py restates the same files generically; the tests pin
 * these loops against it at small sizes, then use the loops where the Python oracle is too slow).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Config 5: u_t = D lap(u) + u (1 - u), order-2 centred 7-point stencil (centered_difference.jl:5-57 with the
 * uniform row [1,-2,1]/h^2, centered_diff_weights.jl:4-40), periodic in x and y (interface_boundary.jl:33-42).
 * One slab of `planes` z planes; the planes below / above it are `lo` / `hi` (NX*NY doubles each): the neighbouring
 * slab's edge plane, or this slab's own opposite plane on a periodic single-slab ring.  x fastest. */
void fisher3d_ref_rhs_slab(double* restrict du, const double* restrict u, const double* restrict lo,
                           const double* restrict hi, int NX, int NY, int planes, double h, double D, int nthreads) {
    const double w = 1.0 / (h * h);
    const size_t P = (size_t)NX * NY;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
    for (int k = 0; k < planes; ++k) {
        const double* um = (k == 0) ? lo : u + (size_t)(k - 1) * P;
        const double* up = (k == planes - 1) ? hi : u + (size_t)(k + 1) * P;
        const double* uc = u + (size_t)k * P;
        double* o = du + (size_t)k * P;
        for (int j = 0; j < NY; ++j) {
            const int jm = (j == 0) ? NY - 1 : j - 1, jp = (j == NY - 1) ? 0 : j + 1;
            for (int i = 0; i < NX; ++i) {
                const int im = (i == 0) ? NX - 1 : i - 1, ip = (i == NX - 1) ? 0 : i + 1;
                const double c = uc[(size_t)j * NX + i];
                const double dxx = w * uc[(size_t)j * NX + im] + -2.0 * w * c + w * uc[(size_t)j * NX + ip];
                const double dyy = w * uc[(size_t)jm * NX + i] + -2.0 * w * c + w * uc[(size_t)jp * NX + i];
                const double dzz = w * um[(size_t)j * NX + i] + -2.0 * w * c + w * up[(size_t)j * NX + i];
                o[(size_t)j * NX + i] = D * (dxx + dyy + dzz) + c * (1.0 - c);
            }
        }
    }
}

/* Config 3: 2-D viscous Burgers, UpwindScheme (order 1), x: Neumann / Neumann, y: Robin / Dirichlet, on uniform or
 * non-uniform grids (examples.burgers_2d):
 *     u_t = -u u_x - v u_y + nu (u_xx + u_yy),   v_t = -u v_x - v v_y + nu (v_xx + v_yy)
 * Per node the reference selects a row of the operator tables (centered_difference.jl:5-57, upwind_difference.jl:131-199)
 * and builds  ifelse(coef > 0, coef * backward row, coef * forward row)  on the cardinalised residual
 * (upwind_difference.jl:192-198); boundary-face nodes are eliminated through the boundary equations
 * (generate_bc_eqs.jl:238-328) with the one-sided rows of the centred first-derivative operator.
 * The ROWS (first tap, weights) come from the Python oracle's table builders (oracle/operators.py + OracleProblem
 * .upwind_row / .centered_row) and are passed in as arrays indexed by the 1-based node number; the loops, the index
 * arithmetic and the boundary elimination below are this file's own.
 *   wm/sm: backward rows (2 weights), wp/sp: forward rows (2), w2/s2: second-derivative rows (3), per dimension;
 *   blo/bhi: one-sided first-derivative rows at the first / last node on nodes 1..3 / n-2..n.
 * State layout: interior nodes 2..n-1 of each dimension, x fastest, u block then v block. */
static void burgers_fill(double* F, const double* un, int nx, int ny, const double* bxlo, const double* bxhi,
                         const double* bylo, double robin_g, const double* top, double top_const) {
    const int mx = nx - 2;
    for (int j = 2; j <= ny - 1; ++j)
        for (int i = 2; i <= nx - 1; ++i) F[(size_t)(j - 1) * nx + (i - 1)] = un[(size_t)(j - 2) * mx + (i - 2)];
#define FF(i, j) F[(size_t)((j) - 1) * nx + ((i) - 1)]
    for (int j = 2; j <= ny - 1; ++j) {                       /* x faces: Dx u = 0 */
        FF(1, j) = -(bxlo[1] * FF(2, j) + bxlo[2] * FF(3, j)) / bxlo[0];
        FF(nx, j) = -(bxhi[0] * FF(nx - 2, j) + bxhi[1] * FF(nx - 1, j)) / bxhi[2];
    }
    for (int i = 2; i <= nx - 1; ++i) {                       /* y = 0: u + u_y / 2 = g;  y = 1: Dirichlet data */
        FF(i, 1) = (robin_g - 0.5 * (bylo[1] * FF(i, 2) + bylo[2] * FF(i, 3))) / (1.0 + 0.5 * bylo[0]);
        FF(i, ny) = top ? top[i - 1] : top_const;
    }
#undef FF
}

void burgers2d_ref_rhs(double* restrict du, const double* restrict u, int nx, int ny,
                       const double* wmx, const int* smx, const double* wpx, const int* spx, const double* w2x, const int* s2x,
                       const double* wmy, const int* smy, const double* wpy, const int* spy, const double* w2y, const int* s2y,
                       const double* bxlo, const double* bxhi, const double* bylo, const double* xg, double nu, double t,
                       double* restrict workU, double* restrict workV, int nthreads) {
    const int mx = nx - 2, my = ny - 2;
    const size_t nint = (size_t)mx * my;
    double* topu = workU + (size_t)nx * ny;                   /* scratch behind the full array: nx doubles */
    for (int i = 1; i <= nx; ++i) topu[i - 1] = 0.2 - 0.5 * sin(M_PI * xg[i - 1]) * exp(-t);
    burgers_fill(workU, u, nx, ny, bxlo, bxhi, bylo, 0.2, topu, 0.0);
    burgers_fill(workV, u + nint, nx, ny, bxlo, bxhi, bylo, -0.1, NULL, -0.1);
#define AT(F, i, j) F[(size_t)((j) - 1) * nx + ((i) - 1)]
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
    for (int j = 2; j <= ny - 1; ++j) {
        for (int i = 2; i <= nx - 1; ++i) {
            const double uc = AT(workU, i, j), vc = AT(workV, i, j);
            double out[2];
            for (int sp = 0; sp < 2; ++sp) {
                const double* F = sp ? workV : workU;
                const double dxm = wmx[2 * i] * AT(F, smx[i], j) + wmx[2 * i + 1] * AT(F, smx[i] + 1, j);
                const double dxp = wpx[2 * i] * AT(F, spx[i], j) + wpx[2 * i + 1] * AT(F, spx[i] + 1, j);
                const double dym = wmy[2 * j] * AT(F, i, smy[j]) + wmy[2 * j + 1] * AT(F, i, smy[j] + 1);
                const double dyp = wpy[2 * j] * AT(F, i, spy[j]) + wpy[2 * j + 1] * AT(F, i, spy[j] + 1);
                const double dxx = w2x[3 * i] * AT(F, s2x[i], j) + w2x[3 * i + 1] * AT(F, s2x[i] + 1, j) + w2x[3 * i + 2] * AT(F, s2x[i] + 2, j);
                const double dyy = w2y[3 * j] * AT(F, i, s2y[j]) + w2y[3 * j + 1] * AT(F, i, s2y[j] + 1) + w2y[3 * j + 2] * AT(F, i, s2y[j] + 2);
                const double ax = (uc > 0.0) ? uc * dxm : uc * dxp;
                const double ay = (vc > 0.0) ? vc * dym : vc * dyp;
                out[sp] = -(ax + ay) + nu * (dxx + dyy);
            }
            du[(size_t)(j - 2) * mx + (i - 2)] = out[0];
            du[nint + (size_t)(j - 2) * mx + (i - 2)] = out[1];
        }
    }
#undef AT
}
