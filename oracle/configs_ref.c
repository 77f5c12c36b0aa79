/* Looped C restatements of the other BASELINE.json configurations at benchmark size — TEST/BASELINE
 * INFRASTRUCTURE ONLY (the product never links or loads this file).
 *
 * Each function restates, per grid point, what the reference's discretisation produces for that
 * problem (the Python oracle oracle/discretize.py restates the same files generically; the tests pin
 * these loops against it at small sizes, then use the loops where the Python oracle is too slow).
 */
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
