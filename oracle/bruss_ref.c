/* CPU restatement of the reference's generated Brusselator RHS — TEST/BASELINE INFRASTRUCTURE ONLY.
 *
 * "Reference-equivalent CPU restatement" (BASELINE.md §3): per grid point exactly the expression
 * the reference generates (docs/src/generated/bruss_code.md:82-113), i.e. the symbolically
 * simplified form with literal coefficients
 *     du = 1.0 + c0u*u + a*(u_E + u_W + u_N + u_S) + u^2 v (+ 5.0 inside the forcing disk, t >= 1.1)
 *     dv = 3.4*u + a*(v_E + v_W + v_N + v_S) + c0v*v - u^2 v
 * with a = alpha/dx^2, c0u = -4a - 4.4, c0v = -4a, looped over the interior instead of unrolled.
 * State layout = the reference's flat unknown vector (x fastest, u block then v block); unknown
 * (i,j), i,j = 0..N-1, is grid node (i+2, j+2) of the N+1 periodic nodes (interior_map.jl:5-9),
 * neighbours wrap modulo N (interface_boundary.jl:33-42).
 *
 * The reference's generated f! is single-threaded; `nthreads` > 1 adds OpenMP over rows so the
 * baseline can also be quoted on all host cores.  Not used by the product path.
 */
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void bruss_ref_rhs(double* restrict du, const double* restrict u, const double* restrict xg,
                   const double* restrict yg, int N, double alpha, double t, int nthreads) {
    const double dx = xg[1] - xg[0];
    const double a = alpha * (1.0 / (dx * dx));
    const double c0u = -4.0 * a - 4.4, c0v = -4.0 * a;
    const double* U = u;
    const double* V = u + (size_t)N * N;
    double* dU = du;
    double* dV = du + (size_t)N * N;
    const int forcing = t >= 1.1;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
    for (int j = 0; j < N; ++j) {
        const int jm = (j == 0) ? N - 1 : j - 1, jp = (j == N - 1) ? 0 : j + 1;
        const double y = yg[j + 1];
        for (int i = 0; i < N; ++i) {
            const int im = (i == 0) ? N - 1 : i - 1, ip = (i == N - 1) ? 0 : i + 1;
            const size_t c = (size_t)j * N + i;
            const double uc = U[c], vc = V[c];
            const double x = xg[i + 1];
            double f = 0.0;
            if (forcing && ((x - 0.3) * (x - 0.3) + (y - 0.6) * (y - 0.6) <= 0.1 * 0.1)) f = 5.0;
            const double un = U[(size_t)jp * N + i], us = U[(size_t)jm * N + i], ue = U[(size_t)j * N + ip],
                         uw = U[(size_t)j * N + im];
            const double vn = V[(size_t)jp * N + i], vs = V[(size_t)jm * N + i], ve = V[(size_t)j * N + ip],
                         vw = V[(size_t)j * N + im];
            const double uuv = uc * uc * vc;
            dU[c] = 1.0 + c0u * uc + a * un + a * us + a * ue + a * uw + uuv + f;
            dV[c] = 3.4 * uc + a * vn + a * vs + a * ve + a * vw + c0v * vc - uuv;
        }
    }
}

/* One rank's slab of the same RHS (slab decomposition along y): rows j = 0..rows-1 of the slab are
 * global unknown rows first_row + j; the rows below / above the slab come from the neighbouring slabs (lo_* / hi_*,
 * NX doubles each; for a periodic ring these are the wrapped rows).  u = [U(rows x NX), V(rows x NX)], x fastest.
 * yg[j] is the y coordinate of slab row j.  Same expression, same operation order as bruss_ref_rhs. */
void bruss_ref_rhs_slab(double* restrict du, const double* restrict u, const double* restrict lo_u,
                        const double* restrict hi_u, const double* restrict lo_v, const double* restrict hi_v,
                        const double* restrict xg, const double* restrict yg, int NX, int rows, double alpha,
                        double t, int nthreads) {
    const double dx = xg[1] - xg[0];
    const double a = alpha * (1.0 / (dx * dx));
    const double c0u = -4.0 * a - 4.4, c0v = -4.0 * a;
    const double* U = u;
    const double* V = u + (size_t)NX * rows;
    double* dU = du;
    double* dV = du + (size_t)NX * rows;
    const int forcing = t >= 1.1;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
    for (int j = 0; j < rows; ++j) {
        const double* Us = (j == 0) ? lo_u : U + (size_t)(j - 1) * NX;
        const double* Un = (j == rows - 1) ? hi_u : U + (size_t)(j + 1) * NX;
        const double* Vs = (j == 0) ? lo_v : V + (size_t)(j - 1) * NX;
        const double* Vn = (j == rows - 1) ? hi_v : V + (size_t)(j + 1) * NX;
        const double y = yg[j];
        for (int i = 0; i < NX; ++i) {
            const int im = (i == 0) ? NX - 1 : i - 1, ip = (i == NX - 1) ? 0 : i + 1;
            const size_t c = (size_t)j * NX + i;
            const double uc = U[c], vc = V[c];
            const double x = xg[i + 1];
            double f = 0.0;
            if (forcing && ((x - 0.3) * (x - 0.3) + (y - 0.6) * (y - 0.6) <= 0.1 * 0.1)) f = 5.0;
            const double un = Un[i], us = Us[i], ue = U[(size_t)j * NX + ip], uw = U[(size_t)j * NX + im];
            const double vn = Vn[i], vs = Vs[i], ve = V[(size_t)j * NX + ip], vw = V[(size_t)j * NX + im];
            const double uuv = uc * uc * vc;
            dU[c] = 1.0 + c0u * uc + a * un + a * us + a * ue + a * uw + uuv + f;
            dV[c] = 3.4 * uc + a * vn + a * vs + a * ve + a * vw + c0v * vc - uuv;
        }
    }
}

int bruss_ref_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
