// Parser for the serialized stencil program (text IR).
//
// The stencil program is what the new lowering pass emits beside
// src/array_discretization.jl: per-variable interior boxes (interior_map.jl:105-115), per-term
// stencil-row tables (centered_difference.jl:16-27, upwind_difference.jl:8-26,
// half_offset_centred_difference.jl:9-69), ghost rules (generate_bc_eqs.jl) and the pointwise
// expressions in RPN.  Numbers are C99 hex floats so weights survive bit-exactly.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <sstream>

#include "mol_internal.h"

namespace mol {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
const char* last_error_cstr() { return g_err.c_str(); }

static bool to_d(const std::string& s, double& v) {
    char* end = nullptr;
    v = strtod(s.c_str(), &end);
    return end && *end == 0 && !s.empty();
}
static bool to_i(const std::string& s, int& v) {
    char* end = nullptr;
    long x = strtol(s.c_str(), &end, 10);
    v = (int)x;
    return end && *end == 0 && !s.empty();
}

#define NEED(cond, msg)                                                                      \
    do {                                                                                     \
        if (!(cond)) return fail(MOL_E_PARSE, std::string("stencil program line ") +         \
                                 std::to_string(lineno) + ": " + (msg));                     \
    } while (0)

// ---- non-uniform WENO5: the u-independent part of the reconstruction, built once per table ---------------------------
// The reference evaluates nonuniform_weno.jl:120-163 per point with numeric geometry; only the u-dependent arithmetic
// has to run per RHS evaluation.  Each sub-stencil interpolant is a quadratic through (a, b, c): with Lagrange's form
//     p'(x)  = ua ((x-b)+(x-c))/((a-b)(a-c)) + ub ((x-a)+(x-c))/((b-a)(b-c)) + uc ((x-a)+(x-b))/((c-a)(c-b)),
//     p''    = 2 ua/((a-b)(a-c)) + 2 ub/((b-a)(b-c)) + 2 uc/((c-a)(c-b)),
// and both rows sum to zero, so they act on the differences (ub-ua), (uc-ub).  The ideal weights d_k make the combination
// of the three sub-stencil derivatives at x_i equal to the five-point Lagrange derivative there; node 1 belongs to
// sub-stencil 0 only and node 5 to sub-stencil 2 only, hence d0 = W5_1 / w^(0)_a and d2 = W5_5 / w^(2)_c.
// Explicit rows (wall targets, irregular charts) get a record of kWenoRec doubles; core rows (centre target on five
// consecutive nodes) share three per-interval arrays, see kernels/mol_device.cuh.
namespace {

// Simpson cell of target T relative to the reconstruction point x_i = x[T-1]: [x_i + sL, x_i + sR] (half cells at the
// walls).  Formed from node spacings, which are exact in floating point for neighbouring nodes, rather than from
// midpoints (x_a + x_b)/2, whose rounding is of relative size eps * |x| / h on a clustered grid.
struct WenoCell { double xi, sL, sR; };

WenoCell weno_cell(const double x[5], int T) {
    const int k = T - 1;
    WenoCell c;
    c.xi = x[k];
    c.sL = (k == 0) ? 0.0 : -0.5 * (x[k] - x[k - 1]);
    c.sR = (k == 4) ? 0.0 : 0.5 * (x[k + 1] - x[k]);
    return c;
}

// weight of node j in the first derivative at xt of the Lagrange interpolant through x[0..n)
double lagrange_d1(const double* x, int n, int j, double xt) {
    double den = 1.0, num = 0.0;
    for (int l = 0; l < n; ++l)
        if (l != j) den *= x[j] - x[l];
    for (int m = 0; m < n; ++m) {
        if (m == j) continue;
        double prod = 1.0;
        for (int l = 0; l < n; ++l)
            if (l != j && l != m) prod *= xt - x[l];
        num += prod;
    }
    return num / den;
}

void weno_record(const double x[5], int T, double* R) {
    const WenoCell c = weno_cell(x, T);
    const double sL = c.sL, sR = c.sR, dx = sR - sL;
    double wa0 = 0.0, wc2 = 0.0;
    for (int k = 0; k < 3; ++k) {
        const double* s = x + k;
        const double wa = lagrange_d1(s, 3, 0, c.xi), wc = lagrange_d1(s, 3, 2, c.xi);
        R[0 + k] = -wa;                                              // r_k = ra_k (ub-ua) + rb_k (uc-ub)
        R[3 + k] = wc;
        R[6 + k] = -2.0 / ((s[0] - s[1]) * (s[0] - s[2]));           // c_k = p_k''
        R[9 + k] = 2.0 / ((s[2] - s[0]) * (s[2] - s[1]));
        if (k == 0) wa0 = wa;
        if (k == 2) wc2 = wc;
    }
    R[12] = dx * dx;
    R[13] = dx * dx * (sL + sR);
    R[14] = dx * dx * (sL * sL + sL * sR + sR * sR) / 3.0 + dx * dx * dx * dx;
    const double d0 = lagrange_d1(x, 5, 0, c.xi) / wa0, d2 = lagrange_d1(x, 5, 4, c.xi) / wc2;
    const double d[3] = {d0, 1.0 - d0 - d2, d2};
    double sp = 0.0, sm = 0.0;
    for (int k = 0; k < 3; ++k) {                                    // Shi-Hu-Shu splitting, theta = 3
        const double dp = 0.5 * (d[k] + 3.0 * std::fabs(d[k])), dm = dp - d[k];
        R[15 + k] = dp;
        R[18 + k] = dm;
        sp += dp;
        sm += dm;
    }
    R[21] = sp;
    R[22] = sm;
}

}  // namespace

bool weno_nu_tables(Program& P, std::string& err) {
    // which (variable, dimension) uses each table, and is it non-uniform (dx == 0 in the token)?
    for (const Rpn& eq : P.eqs)
        for (const std::string& tk : eq) {
            if (tk.size() < 2 || tk[0] != 'W' || tk[1] != ':') continue;
            int id = 0, var = 0, dim = 0;
            char ebuf[64] = {0}, dbuf[64] = {0};
            if (sscanf(tk.c_str(), "W:%d:%d:%d:%63[^:]:%63s", &id, &var, &dim, ebuf, dbuf) != 5) continue;
            auto it = P.wtabs.find(id);
            if (it == P.wtabs.end()) { err = "W token references an unknown wtab"; return false; }
            if (var < 0 || var >= P.nvar || dim < 0 || dim >= P.ndim) { err = "W token: bad variable / dimension"; return false; }
            WTab& T = it->second;
            T.var = var;
            T.dim = dim;
            T.nu = strtod(dbuf, nullptr) == 0.0;
        }
    for (auto& kv : P.wtabs) {
        WTab& T = kv.second;
        if (!T.nu) continue;
        const Grid& G = P.grid[T.dim];
        const int n = G.n;
        const bool per = P.vars[T.var].per[T.dim] != 0;
        const double period = G.x[n - 1] - G.x[0];
        // chart coordinate of a tap (1-based node number, possibly past a periodic seam)
        auto X = [&](int j) -> double {
            double shift = 0.0;
            if (per) {
                if (j <= 1) { j += n - 1; shift = -period; }
                else if (j > n) { j -= n - 1; shift = period; }
            }
            j = std::max(1, std::min(n, j));
            return G.x[j - 1] + shift;
        };
        if (T.has_core) {
            T.glo = T.core_lo - 2;
            T.glen = T.core_hi - T.core_lo + 4;
            T.goff = (int)P.tabw.size();
            P.tabw.resize(P.tabw.size() + (size_t)3 * T.glen, 1.0);
            double* g = P.tabw.data() + T.goff;
            // centre-target ideal weights at every core node (same expressions as the kernel): all positive?
            T.allpos = true;
            for (int idx = T.core_lo; idx <= T.core_hi && T.allpos; ++idx) {
                const double ha = X(idx - 1) - X(idx - 2), hb = X(idx) - X(idx - 1), hc = X(idx + 1) - X(idx), hd = X(idx + 2) - X(idx + 1);
                const double s3a = ha + hb + hc, s4 = s3a + hd, s3b = hb + hc + hd;
                const double den = s3a * s4 * s3b, d0 = hc * (hc + hd) * s3b, d2 = (ha + hb) * hb * s3a;
                if (!(d0 > 0.0 && d2 > 0.0 && den - d0 - d2 > 1e-9 * den)) T.allpos = false;
            }
            for (int q = 0; q < T.glen; ++q) {
                const int j = T.glo + q;
                const double h = X(j + 1) - X(j), h2 = X(j + 2) - X(j);
                if (!(h > 0.0) || !(h2 > 0.0)) { err = "WENO needs strictly increasing grid coordinates"; return false; }
                g[q] = h;
                g[T.glen + q] = 1.0 / h;
                g[2 * T.glen + q] = 1.0 / h2;
            }
        }
        T.roff = (int)P.tabw.size();
        T.nrec = 0;
        for (int r = 0; r < T.nrows; ++r) {
            const int idx = T.first + r;
            if (T.has_core && idx >= T.core_lo && idx <= T.core_hi) continue;
            double x[5];
            for (int k = 0; k < 5; ++k) x[k] = X(T.start[r] + k);
            for (int k = 0; k < 4; ++k)
                if (!(x[k + 1] > x[k])) { err = "WENO needs strictly increasing grid coordinates"; return false; }
            if (T.target[r] < 1 || T.target[r] > 5) { err = "WENO target outside 1..5"; return false; }
            P.tabw.resize(P.tabw.size() + kWenoRec);
            weno_record(x, T.target[r], P.tabw.data() + T.roff + (size_t)T.nrec * kWenoRec);
            ++T.nrec;
        }
    }
    return true;
}

// ---- packed per-node records for the tiled kernel on non-uniform axes --------------------------------------------------
// On a non-uniform axis every core node has its own row of weights per operator ("shape core" tables: same taps, per-node
// weights), and non-uniform WENO5 needs the spacings around the node and their reciprocals.  Read one by one from the
// tables that is 12-14 dependent global loads per node, and the kernels wait on them (ncu: long-scoreboard stalls; 32 %
// of the HBM roofline for the 2-D upwind + diffusion program, 23 % for 1-D WENO5).  Here everything a node needs along
// one dimension is packed into ONE record (16-byte aligned, shared by every variable that uses the same tables); the
// tiled kernel copies the records of a tile's columns and rows into shared memory together with the tile (cp.async,
// same commit group) and the arithmetic reads them from there (the row records broadcast).
//   record of node j = [weights of every shape-core table of the dimension at row j] ++
//                      [h_j, 1/h_j, 1/(x_{j+2} - x_j) of interval j, per non-uniform WENO table of the dimension]
// A WENO evaluation at node i reads the records i-2 .. i+1, hence 2 / 1 records of halo around a tile.  The array starts
// wrec_hl records before the core box and ends kWrecPad records after it (zeros: overhanging tiles stay inside it).
constexpr int kWrecPad = 4096;

void weight_records(Program& P) {
    if (!P.has_core || P.ndim > 2) return;
    for (int d = 0; d < P.ndim; ++d) {
        std::vector<const Tab*> tabs;
        std::vector<const WTab*> wtabs;
        for (const Rpn& eq : P.eqs)
            for (const std::string& tk : eq) {
                int id = 0, var = 0, dim = 0;
                if (tk.size() < 2 || tk[1] != ':') continue;
                if (tk[0] == 'L' && sscanf(tk.c_str(), "L:%d:%d:%d", &id, &var, &dim) == 3 && dim == d) {
                    auto it = P.tabs.find(id);
                    if (it == P.tabs.end()) continue;
                    const Tab& T = it->second;
                    const bool literal = T.has_core && T.core_lo <= P.clo[d] && T.core_hi >= P.chi[d];
                    if (literal || !T.has_score || T.score_lo > P.clo[d] || T.score_hi < P.chi[d]) continue;
                    if (!P.wrec_pos.count(id)) { P.wrec_pos[id] = -1; tabs.push_back(&T); }
                } else if (getenv("MOL_WENO_STAGE") && tk[0] == 'W' && sscanf(tk.c_str(), "W:%d:%d:%d", &id, &var, &dim) == 3 && dim == d) {
                    // WENO geometry in the staged records: measured NOT to pay (B200: 1-D 2^22 nodes 47.7 us staged vs
                    // 45.0 us through the read-only path; 2-D 2048^2 64.3 vs 62.4 us) -- that kernel is bound by issue
                    // slots, not by the latency of its 11 geometry loads.  Kept behind MOL_WENO_STAGE for experiments.
                    auto it = P.wtabs.find(id);
                    if (it == P.wtabs.end()) continue;
                    const WTab& T = it->second;
                    if (!T.nu || !T.has_core || T.core_lo > P.clo[d] || T.core_hi < P.chi[d]) continue;
                    if (!P.wrec_wpos.count(id)) { P.wrec_wpos[id] = -1; wtabs.push_back(&T); }
                }
            }
        if (tabs.empty() && wtabs.empty()) continue;
        int stride = 0;
        for (const Tab* T : tabs) { P.wrec_pos[T->id] = stride; stride += T->score_n; }
        for (const WTab* T : wtabs) { P.wrec_wpos[T->id] = stride; stride += 3; }
        stride = (stride + 1) / 2 * 2;
        const int hl = wtabs.empty() ? 0 : 2, hh = wtabs.empty() ? 0 : 1;
        const int ncore = P.chi[d] - P.clo[d] + 1;
        if (P.tabw.size() % 2) P.tabw.push_back(0.0);          // 16-byte alignment of the records
        P.wrec_off[d] = (int)P.tabw.size();
        P.wrec_stride[d] = stride;
        P.wrec_hl[d] = hl;
        P.wrec_hh[d] = hh;
        P.wrec_lo[d] = P.clo[d] - hl;
        P.wrec_n[d] = (hl + ncore + kWrecPad + 1) / 2 * 2;
        P.tabw.resize(P.tabw.size() + (size_t)P.wrec_n[d] * stride, 0.0);
        double* R = P.tabw.data() + P.wrec_off[d];
        // Dimension 0 is stored FIELD-major (field f of node j at R[f * wrec_n + j]): the threads of a warp own
        // consecutive x nodes and read the same field, which must fall into consecutive shared-memory words (node-major
        // records put a whole warp on one bank: measured 34.8 M bank conflicts per sweep).  The other dimensions are
        // node-major: a warp reads ONE record there, which broadcasts.
        const size_t nrec = (size_t)P.wrec_n[d];
        auto at = [&](int node_rel, int field) -> double& {
            return d == 0 ? R[(size_t)field * nrec + node_rel] : R[(size_t)node_rel * stride + field];
        };
        for (int q = 0; q < ncore; ++q)
            for (const Tab* T : tabs) {
                const Row& row = T->rows[P.clo[d] + q - T->first];
                for (int k = 0; k < T->score_n && k < (int)row.w.size(); ++k) at(hl + q, P.wrec_pos[T->id] + k) = row.w[k];
            }
        for (const WTab* T : wtabs) {
            // the table's own per-interval arrays (weno_nu_tables) cover intervals core_lo - 2 .. core_hi + 1
            const double* g = P.tabw.data() + T->goff;
            for (int j = P.clo[d] - 2; j <= P.chi[d] + 1; ++j) {
                const int qg = j - T->glo;
                if (qg < 0 || qg >= T->glen) continue;
                const int pos = P.wrec_wpos[T->id];
                at(j - P.wrec_lo[d], pos) = g[qg];
                at(j - P.wrec_lo[d], pos + 1) = g[T->glen + qg];
                at(j - P.wrec_lo[d], pos + 2) = g[2 * T->glen + qg];
            }
        }
    }
}

int parse_program(const char* text, size_t nbytes, Program& P) {
    std::string all(text, nbytes);
    std::istringstream is(all);
    std::string line;
    int lineno = 0;
    bool header = false, ended = false;
    while (std::getline(is, line)) {
        ++lineno;
        std::istringstream ls(line);
        std::vector<std::string> tk;
        std::string w;
        while (ls >> w) tk.push_back(w);
        if (tk.empty() || tk[0][0] == '#') continue;
        const std::string& k = tk[0];
        auto I = [&](size_t i, int& v) { return i < tk.size() && to_i(tk[i], v); };
        auto D = [&](size_t i, double& v) { return i < tk.size() && to_d(tk[i], v); };
        if (k == "MOLPROG") {
            int ver = 0;
            NEED(I(1, ver) && ver == 1, "unsupported program version");
            header = true;
        } else if (k == "ndim") {
            NEED(I(1, P.ndim) && P.ndim >= 1 && P.ndim <= 3, "ndim must be 1..3");
        } else if (k == "nvar") {
            NEED(I(1, P.nvar) && P.nvar >= 1 && P.nvar <= 8, "nvar must be 1..8");
            P.vars.resize(P.nvar);
            P.eqs.resize(P.nvar);
        } else if (k == "nparam") {
            NEED(I(1, P.nparam) && P.nparam >= 0 && P.nparam <= 64, "nparam must be 0..64");
            P.pname.resize(P.nparam);
            P.pdefault.assign(P.nparam, 0.0);
        } else if (k == "param") {
            int i;
            NEED(I(1, i) && i >= 0 && i < P.nparam && tk.size() >= 4, "bad param");
            P.pname[i] = tk[2];
            NEED(D(3, P.pdefault[i]), "bad param value");
        } else if (k == "grid") {
            int j;
            NEED(I(1, j) && j >= 0 && j < P.ndim && tk.size() >= 5, "bad grid");
            NEED(I(2, P.grid[j].n) && P.grid[j].n >= 2, "bad grid size");
            P.grid[j].uniform = (tk[3] == "U");
            NEED(D(4, P.grid[j].dx), "bad grid dx");
        } else if (k == "coords") {
            int j;
            NEED(I(1, j) && j >= 0 && j < P.ndim, "bad coords");
            NEED((int)tk.size() == 2 + P.grid[j].n, "coords count does not match grid size");
            P.grid[j].x.resize(P.grid[j].n);
            for (int i = 0; i < P.grid[j].n; ++i) NEED(D(2 + i, P.grid[j].x[i]), "bad coordinate");
        } else if (k == "var") {
            int v;
            NEED(I(1, v) && v >= 0 && v < P.nvar && tk.size() >= 3, "bad var");
            P.vars[v].name = tk[2];
        } else if (k == "interior") {
            int v;
            NEED(I(1, v) && v >= 0 && v < P.nvar && (int)tk.size() == 2 + 2 * P.ndim, "bad interior");
            for (int j = 0; j < P.ndim; ++j) {
                NEED(I(2 + j, P.vars[v].ilo[j]) && I(2 + P.ndim + j, P.vars[v].ihi[j]), "bad interior bounds");
                NEED(P.vars[v].ilo[j] >= 1 && P.vars[v].ihi[j] <= P.grid[j].n && P.vars[v].ilo[j] <= P.vars[v].ihi[j],
                     "interior box outside the grid");
            }
        } else if (k == "periodic") {
            int v;
            NEED(I(1, v) && v >= 0 && v < P.nvar && (int)tk.size() == 2 + P.ndim, "bad periodic");
            for (int j = 0; j < P.ndim; ++j) NEED(I(2 + j, P.vars[v].per[j]), "bad periodic flag");
        } else if (k == "tab") {
            Tab T;
            NEED(I(1, T.id) && I(2, T.L) && I(3, T.nrows) && I(4, T.first), "bad tab");
            NEED(T.L >= 1 && T.L <= 16 && T.nrows >= 1, "bad tab shape");
            T.rows.resize(T.nrows);
            T.have.assign(T.nrows, 0);
            P.tabs[T.id] = T;
        } else if (k == "core") {
            int id;
            NEED(I(1, id) && P.tabs.count(id), "core: unknown tab");
            Tab& T = P.tabs[id];
            NEED(I(2, T.core_lo) && I(3, T.core_hi) && I(4, T.core_off) && (int)tk.size() == 5 + T.L, "bad core");
            T.core_w.resize(T.L);
            for (int q = 0; q < T.L; ++q) NEED(D(5 + q, T.core_w[q]), "bad core weight");
            T.has_core = T.core_hi >= T.core_lo;
            // The literal row is padded with zeros to the table's row length L (one-sided boundary rows can be longer
            // than interior ones, e.g. the outer operator of the nonlinear Laplacian: 2 interior taps, 3 at the wall).
            // The table-driven kernel walks `n` taps per row: a padded tap would make it evaluate a node -- or, for
            // half-point tables, a whole inner row -- past the end of the grid and multiply it by 0.0 (NaN if that
            // garbage is not finite).  The reference drops zero-weight terms symbolically; so do the expanded rows.
            std::vector<double> cw = T.core_w;
            while (!cw.empty() && cw.back() == 0.0) cw.pop_back();
            for (int idx = T.core_lo; idx <= T.core_hi; ++idx) {
                int r = idx - T.first;
                NEED(r >= 0 && r < T.nrows, "core range outside tab");
                T.rows[r].start = idx + T.core_off;
                T.rows[r].w = cw;
                T.have[r] = 1;
            }
        } else if (k == "score") {
            int id;
            NEED(I(1, id) && P.tabs.count(id), "score: unknown tab");
            Tab& T = P.tabs[id];
            NEED((int)tk.size() == 6 && I(2, T.score_lo) && I(3, T.score_hi) && I(4, T.score_off) && I(5, T.score_n) &&
                     T.score_n >= 0 && T.score_n <= T.L,
                 "bad score");
            T.has_score = T.score_hi >= T.score_lo;
        } else if (k == "row") {
            int id, idx, nt;
            NEED(I(1, id) && P.tabs.count(id), "row: unknown tab");
            Tab& T = P.tabs[id];
            NEED(I(2, idx), "bad row index");
            int r = idx - T.first;
            NEED(r >= 0 && r < T.nrows, "row outside tab");
            NEED(I(3, T.rows[r].start) && I(4, nt) && nt >= 0 && nt <= T.L && (int)tk.size() == 5 + nt, "bad row");
            T.rows[r].w.resize(nt);
            for (int q = 0; q < nt; ++q) NEED(D(5 + q, T.rows[r].w[q]), "bad row weight");
            T.have[r] = 1;
        } else if (k == "wtab") {
            WTab T;
            NEED(I(1, T.id) && I(2, T.nrows) && I(3, T.first) && T.nrows >= 1, "bad wtab");
            T.start.assign(T.nrows, 0);
            T.target.assign(T.nrows, 3);
            T.have.assign(T.nrows, 0);
            P.wtabs[T.id] = T;
        } else if (k == "wcore") {
            int id;
            NEED(I(1, id) && P.wtabs.count(id), "wcore: unknown wtab");
            WTab& T = P.wtabs[id];
            NEED(I(2, T.core_lo) && I(3, T.core_hi), "bad wcore");
            T.has_core = T.core_hi >= T.core_lo;
            for (int idx = T.core_lo; idx <= T.core_hi; ++idx) {
                int r = idx - T.first;
                NEED(r >= 0 && r < T.nrows, "wcore range outside wtab");
                T.start[r] = idx - 2;
                T.target[r] = 3;
                T.have[r] = 1;
            }
        } else if (k == "wrow") {
            int id, idx;
            NEED(I(1, id) && P.wtabs.count(id), "wrow: unknown wtab");
            WTab& T = P.wtabs[id];
            NEED(I(2, idx), "bad wrow");
            int r = idx - T.first;
            NEED(r >= 0 && r < T.nrows, "wrow outside wtab");
            NEED(I(3, T.start[r]) && I(4, T.target[r]), "bad wrow");
            T.have[r] = 1;
        } else if (k == "fn") {
            int id, nt;
            NEED(I(1, id) && I(2, nt) && (int)tk.size() == 3 + nt, "bad fn");
            P.fns[id] = Rpn(tk.begin() + 3, tk.end());
        } else if (k == "ghost") {
            Ghost g;
            int ntaps;
            NEED(I(1, g.var) && I(2, g.dim) && I(3, g.node) && I(4, ntaps), "bad ghost");
            NEED(g.var >= 0 && g.var < P.nvar && g.dim >= 0 && g.dim < P.ndim && ntaps >= 0, "bad ghost ids");
            size_t pos = 5;
            for (int q = 0; q < ntaps; ++q) {
                GhostTap tp;
                NEED(I(pos, tp.var) && I(pos + 1, tp.node) && D(pos + 2, tp.coef), "bad ghost tap");
                NEED(tp.var >= 0 && tp.var < P.nvar, "bad ghost tap variable");
                g.taps.push_back(tp);
                pos += 3;
            }
            int nt;
            NEED(I(pos, nt) && tk.size() == pos + 1 + nt, "bad ghost expression");
            g.expr = Rpn(tk.begin() + pos + 1, tk.end());
            P.ghosts.push_back(g);
        } else if (k == "ghostx") {
            // ghost rule whose tap coefficients are expressions: ghostx v dim node ntaps nG G.. (var node nc coef..)*
            Ghost g;
            int ntaps, ng;
            NEED(I(1, g.var) && I(2, g.dim) && I(3, g.node) && I(4, ntaps) && I(5, ng), "bad ghostx");
            NEED(g.var >= 0 && g.var < P.nvar && g.dim >= 0 && g.dim < P.ndim && ntaps >= 0 && ng >= 1, "bad ghostx ids");
            size_t pos = 6;
            NEED(tk.size() >= pos + ng, "bad ghostx expression");
            g.expr = Rpn(tk.begin() + pos, tk.begin() + pos + ng);
            pos += ng;
            for (int q = 0; q < ntaps; ++q) {
                GhostTap tp;
                int nc;
                NEED(I(pos, tp.var) && I(pos + 1, tp.node) && I(pos + 2, nc) && nc >= 1 && tk.size() >= pos + 3 + nc, "bad ghostx tap");
                NEED(tp.var >= 0 && tp.var < P.nvar, "bad ghostx tap variable");
                tp.coef = 0.0;
                g.taps.push_back(tp);
                g.tapexpr.push_back(Rpn(tk.begin() + pos + 3, tk.begin() + pos + 3 + nc));
                pos += 3 + nc;
            }
            NEED(pos == tk.size(), "trailing tokens in ghostx");
            P.ghosts.push_back(g);
        } else if (k == "eq") {
            int v, nt;
            NEED(I(1, v) && v >= 0 && v < P.nvar && I(2, nt) && (int)tk.size() == 3 + nt, "bad eq");
            P.eqs[v] = Rpn(tk.begin() + 3, tk.end());
        } else if (k == "corebox") {
            NEED((int)tk.size() == 1 + 2 * P.ndim, "bad corebox");
            for (int j = 0; j < P.ndim; ++j) NEED(I(1 + j, P.clo[j]) && I(1 + P.ndim + j, P.chi[j]), "bad corebox");
            P.has_core = true;
            for (int j = 0; j < P.ndim; ++j)
                if (P.chi[j] < P.clo[j]) P.has_core = false;
        } else if (k == "end") {
            ended = true;
            break;
        } else {
            NEED(false, "unknown directive '" + k + "'");
        }
    }
    NEED(header && ended, "missing MOLPROG header or end");
    NEED(P.ndim > 0 && P.nvar > 0, "ndim/nvar missing");
    for (int j = 0; j < P.ndim; ++j) NEED((int)P.grid[j].x.size() == P.grid[j].n, "grid coordinates missing");
    for (int v = 0; v < P.nvar; ++v) NEED(!P.eqs[v].empty(), "equation missing for a variable");
    // state layout: variable-major, first spatial index fastest
    int64_t off = 0;
    for (int v = 0; v < P.nvar; ++v) {
        P.voff[v] = off;
        int64_t sz = 1;
        for (int j = 0; j < P.ndim; ++j) sz *= P.vars[v].ext(j);
        off += sz;
    }
    P.nstate = off;
    // flatten tables
    for (auto& kv : P.tabs) {
        Tab& T = kv.second;
        for (int r = 0; r < T.nrows; ++r)
            NEED(T.have[r], "tab " + std::to_string(T.id) + " has an undefined row");
        T.woff = (int)P.tabw.size();
        T.soff = (int)P.tabs_flat.size();
        for (int r = 0; r < T.nrows; ++r) {
            for (int q = 0; q < T.L; ++q) P.tabw.push_back(q < (int)T.rows[r].w.size() ? T.rows[r].w[q] : 0.0);
            P.tabs_flat.push_back(T.rows[r].start);
            P.tabs_flat.push_back((int)T.rows[r].w.size());
        }
    }
    for (auto& kv : P.wtabs) {
        WTab& T = kv.second;
        for (int r = 0; r < T.nrows; ++r)
            NEED(T.have[r], "wtab " + std::to_string(T.id) + " has an undefined row");
    }
    {
        std::string werr;
        NEED(weno_nu_tables(P, werr), werr);
    }
    for (auto& kv : P.wtabs) {
        WTab& T = kv.second;
        T.soff = (int)P.tabs_flat.size();
        int rec = 0;
        for (int r = 0; r < T.nrows; ++r) {
            const int idx = T.first + r;
            const bool core = T.has_core && idx >= T.core_lo && idx <= T.core_hi;
            P.tabs_flat.push_back(T.start[r]);
            // target in the low three bits; explicit rows of a non-uniform table also carry 1 + their record number
            P.tabs_flat.push_back(T.target[r] | ((T.nu && !core) ? (++rec << 3) : 0));
        }
    }
    weight_records(P);
    if (P.tabw.empty()) P.tabw.push_back(0.0);
    if (P.tabs_flat.empty()) P.tabs_flat.push_back(0);
    return MOL_OK;
}

}  // namespace mol
