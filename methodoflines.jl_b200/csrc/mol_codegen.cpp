// Code generation: stencil program -> CUDA source for NVRTC.
//
// Mirrors what the reference does with RuntimeGeneratedFunctions (the generated f! of
// docs/src/generated/bruss_code.md:46-118 has literal weights baked in): literal stencil weights
// are emitted for the tiled core kernel, table lookups for the generic kernel; the pointwise
// expressions (reaction terms, upwind ifelse, nonlinear-Laplacian coefficients, boundary data)
// are emitted as straight-line SSA code from their RPN form.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <sstream>

#include "mol_internal.h"

namespace mol {

static std::string hexd(double v) {
    if (std::isnan(v)) return "(0.0/0.0)";
    if (std::isinf(v)) return v > 0 ? "(1.0/0.0)" : "(-1.0/0.0)";
    char buf[64];
    snprintf(buf, sizeof buf, "%a", v);
    std::string s(buf);
    if (s[0] == '-') return "(" + s + ")";
    return s;
}

enum Mode { GENERIC, TILE, FN, GHOST };

struct Val {
    std::string s;
    bool is_bool;
};

struct Emitter {
    const Program& P;
    Mode mode;
    std::ostringstream code;
    int tmp = 0;
    std::string err;
    // tile bookkeeping: max reach per dim
    int* reach;   // int[3]
    std::map<int, std::string> ucache;

    // dual = true: the same expressions on dual numbers (value + tangent; kernels/mol_jvp.cuh) for the
    // Jacobian-vector product: the table-driven forms (GENERIC / FN / GHOST modes) and the tiled form (TILE: MOL_SD pairs
    // the cell of the u tile with the same cell of the v tile).
    bool dual = false;

    Emitter(const Program& p, Mode m, int* r, bool d = false) : P(p), mode(m), reach(r), dual(d) {}

    const char* num() const { return dual ? "MolDual " : "double "; }
    // state access prefix of the table-driven helpers: (in, c, ...) or (in, jv, c, ...) on dual numbers
    const char* ctx() const { return dual ? "(in, jv, c, " : "(in, c, "; }

    std::string fresh(const std::string& expr, bool b = false) {
        std::string n = "t" + std::to_string(tmp++);
        code << "    const " << (b ? "bool " : num()) << n << " = " << expr << ";\n";
        return n;
    }
    std::string S(int var, int dim, int off) {
        int d[3] = {0, 0, 0};
        d[dim] = off;
        reach[dim] = std::max(reach[dim], std::abs(off));
        std::ostringstream o;
        o << (dual ? "MOL_SD(" : "MOL_S(") << var << "," << d[0] << "," << d[1] << "," << d[2] << ")";
        return o.str();
    }
    bool uses_coord[3] = {false, false, false};
    std::string coord(int j) {
        if (mode == FN) return "xh" + std::to_string(j);
        if (mode == TILE) { uses_coord[j] = true; return "xc" + std::to_string(j); }
        // (clamped: the tiled kernels also resolve halo cells outside the grid in two dimensions at once -- corner cells no
        // stencil taps -- and a ghost rule evaluated there must keep its coordinate reads inside the array)
        const std::string J = std::to_string(j);
        return "__ldg(c.grid[" + J + "] + min(max(i" + J + ", 1), MOL_N" + J + ") - 1)";
    }
    std::string asd(const Val& v) { return v.is_bool ? "(" + v.s + " ? 1.0 : 0.0)" : v.s; }
    std::string asb(const Val& v) { return v.is_bool ? v.s : "(" + v.s + " != 0.0)"; }

    static std::vector<std::string> split(const std::string& s, char sep) {
        std::vector<std::string> out;
        std::string cur;
        for (char ch : s) {
            if (ch == sep) { out.push_back(cur); cur.clear(); } else cur += ch;
        }
        out.push_back(cur);
        return out;
    }

    bool tile_row(const Tab& T, int dim, std::vector<double>& w, int& off) {
        if (!T.has_core) return false;
        if (T.core_lo > P.clo[dim] || T.core_hi < P.chi[dim]) return false;
        w = T.core_w;
        off = T.core_off;
        return true;
    }

    // linear row op in the current mode; `rel` shifts the row index (half-point tables)
    bool emit_lin(int tabid, int var, int dim, Val& out) {
        auto it = P.tabs.find(tabid);
        if (it == P.tabs.end()) { err = "unknown tab " + std::to_string(tabid); return false; }
        const Tab& T = it->second;
        if (mode == TILE) {
            std::vector<double> w;
            int off;
            std::ostringstream o;
            bool first = true;
            if (tile_row(T, dim, w, off)) {                 // literal weights (uniform grid)
                for (int q = 0; q < T.L; ++q) {
                    if (w[q] == 0.0) continue;
                    if (!first) o << " + ";
                    o << hexd(w[q]) << " * " << S(var, dim, off + q);
                    first = false;
                }
            } else if (T.has_score && T.score_lo <= P.clo[dim] && T.score_hi >= P.chi[dim]) {
                // same taps at every node of the core box, weights of the node's own row from the table
                // (non-uniform grid): consecutive nodes of a warp read consecutive rows, the other dimensions broadcast
                std::ostringstream base;
                // (row index clamped: overhanging tile cells are evaluated but never stored)
                base << "(c.tabw + " << T.woff << " + (mol_i64)min(i" << dim << " - " << T.first << ", " << (T.nrows - 1) << ") * "
                     << T.L << ")";
                const std::string wp = "w" + std::to_string(tmp++);
                // staged flavours: the node's packed record in shared memory (weight_records); otherwise the table row
                auto rp = P.wrec_pos.find(T.id);
                const bool staged = rp != P.wrec_pos.end() && rp->second >= 0 && dim < 2;
                if (staged) {
                    code << "#if MOL_WSTAGE\n    const double* " << wp << " = " << (dim == 0 ? "MOL_WX(0, " : "MOL_WY(0, ") << rp->second
                         << ");\n#else\n    const double* " << wp << " = " << base.str() << ";\n#endif\n";
                } else {
                    code << "    const double* " << wp << " = " << base.str() << ";\n";
                }
                // (staged x records are field-major: consecutive weights of a row are MOL_WFS0 doubles apart)
                const std::string fs = staged ? std::string(dim == 0 ? "MOL_WFS0" : "MOL_WFS1") : "1";
                if (staged) code << "#if MOL_WSTAGE\n#define MOL_FS_" << wp << " " << fs << "\n#else\n#define MOL_FS_" << wp << " 1\n#endif\n";
                for (int q = 0; q < T.score_n; ++q) {
                    if (!first) o << " + ";
                    if (staged) o << "mol_gld<(MOL_WSTAGE != 0)>(" << wp << " + " << q << " * MOL_FS_" << wp << ")";
                    else o << "__ldg(" << wp << " + " << q << ")";
                    o << " * " << S(var, dim, T.score_off + q);
                    first = false;
                }
            } else {
                err = "tab has no core row covering the core box";
                return false;
            }
            if (first) o << "0.0";
            out = {fresh(o.str()), false};
        } else {
            std::ostringstream o;
            o << (dual ? "mol_lin_d<" : "mol_lin_g<") << var << "," << dim << ">" << ctx() << T.woff << ", " << T.soff << ", " << T.L
              << ", i" << dim << " - " << T.first << ", i0, i1, i2)";
            out = {fresh(o.str()), false};
        }
        return true;
    }

    // mixed derivative Dx Dy u: product of the two centred first-derivative rows (2nd_order_mixed_deriv.jl:5-22)
    bool emit_mixed(const std::vector<std::string>& f, Val& out) {
        int tx = atoi(f[1].c_str()), ty = atoi(f[2].c_str()), var = atoi(f[3].c_str());
        int dx = atoi(f[4].c_str()), dy = atoi(f[5].c_str());
        if (!P.tabs.count(tx) || !P.tabs.count(ty) || var < 0 || var >= P.nvar || dx < 0 || dy < 0 || dx >= P.ndim ||
            dy >= P.ndim || dx == dy) {
            err = "mixed derivative references an unknown table / variable / dimension";
            return false;
        }
        if (mode == TILE) { err = "mixed derivatives run through the table-driven kernel"; return false; }
        const Tab &TX = P.tabs.at(tx), &TY = P.tabs.at(ty);
        std::ostringstream o;
        o << (dual ? "mol_mixed_d<" : "mol_mixed_g<") << var << "," << dx << "," << dy << ">" << ctx() << TX.woff << ", " << TX.soff
          << ", " << TX.L << ", i" << dx << " - " << TX.first << ", " << TY.woff << ", " << TY.soff << ", " << TY.L << ", i" << dy
          << " - " << TY.first << ", i0, i1, i2)";
        out = {fresh(o.str()), false};
        return true;
    }

    bool emit_weno(const std::vector<std::string>& f, Val& out) {
        int id = atoi(f[1].c_str()), var = atoi(f[2].c_str()), dim = atoi(f[3].c_str());
        double eps = strtod(f[4].c_str(), nullptr), dx = strtod(f[5].c_str(), nullptr);
        auto it = P.wtabs.find(id);
        if (it == P.wtabs.end()) { err = "unknown wtab"; return false; }
        const WTab& T = it->second;
        if (mode == TILE) {
            if (!(T.has_core && T.core_lo <= P.clo[dim] && T.core_hi >= P.chi[dim])) {
                err = "WENO table has no core covering the core box";
                return false;
            }
            std::ostringstream o;
            if (dx != 0.0) {
                o << (dual ? "mol_weno5_uniform_d(" : "mol_weno5_uniform(") << S(var, dim, -2) << ", " << S(var, dim, -1) << ", " << S(var, dim, 0) << ", "
                  << S(var, dim, 1) << ", " << S(var, dim, 2) << ", " << hexd(eps) << ", " << hexd(dx) << ")";
            } else {
                // non-uniform core row: the node's four spacings and their reciprocals from the table's per-interval arrays
                // (index clamped: overhanging tile cells are evaluated but never stored)
                const std::string us = S(var, dim, -2) + ", " + S(var, dim, -1) + ", " + S(var, dim, 0) + ", " + S(var, dim, 1) + ", " +
                                       S(var, dim, 2);
                // (dual numbers: the general form, as in mol_weno_d)
                const std::string pos = dual ? "MolDual, false" : (T.allpos ? "double, true" : "double, false");
                std::ostringstream glob;
                glob << "mol_weno5_nu_core<" << pos << ", false>(" << us << ", c.tabw + " << T.goff << " + (min(i" << dim << ", "
                     << T.core_hi << ") - " << (T.glo + 2) << "), 1, " << T.glen << ", " << hexd(eps) << ")";
                auto rp = P.wrec_wpos.find(T.id);
                if (rp != P.wrec_wpos.end() && rp->second >= 0 && dim < 2) {
                    const std::string r = "t" + std::to_string(tmp++);
                    code << "#if MOL_WSTAGE\n    const " << num() << r << " = mol_weno5_nu_core<" << pos << ", true>(" << us << ", "
                         << (dim == 0 ? "MOL_WX(-2, " : "MOL_WY(-2, ") << rp->second << "), " << (dim == 0 ? "1, MOL_WFS0" : "MOL_WRS1, MOL_WFS1")
                         << ", " << hexd(eps) << ");\n#else\n    const " << num() << r << " = " << glob.str() << ";\n#endif\n";
                    out = {r, false};
                    return true;
                }
                o << glob.str();
            }
            out = {fresh(o.str()), false};
        } else {
            std::ostringstream o;
            o << (dual ? "mol_weno_d<" : "mol_weno_g<") << var << "," << dim << ">" << ctx() << T.soff << ", i" << dim << " - "
              << T.first << ", "
              << hexd(eps) << ", " << hexd(dx) << ", " << T.goff << ", " << T.glo << ", " << T.glen << ", " << T.roff
              << ", " << (T.allpos ? 1 : 0) << ", i0, i1, i2)";
            out = {fresh(o.str()), false};
        }
        return true;
    }

    // nonlinear Laplacian: sum_m wo_m * a(u~_m, x~_m) * (D u)_m   (nonlinear_laplacian.jl:28-103)
    bool emit_nll(const std::vector<std::string>& f, Val& out) {
        int var = atoi(f[1].c_str()), dim = atoi(f[2].c_str()), fn = atoi(f[3].c_str());
        int itab = atoi(f[4].c_str()), dtab = atoi(f[5].c_str()), otab = atoi(f[6].c_str());
        if (!P.tabs.count(itab) || !P.tabs.count(dtab) || !P.tabs.count(otab) || !P.fns.count(fn)) {
            err = "nonlinear Laplacian references unknown table/function";
            return false;
        }
        const Tab &TI = P.tabs.at(itab), &TD = P.tabs.at(dtab), &TO = P.tabs.at(otab);
        std::string r = "t" + std::to_string(tmp++);
        auto fncall = [&](const std::string& uh, const std::string& xh) {
            std::ostringstream o;
            o << (dual ? "mol_fnd_" : "mol_fn_") << fn << "(" << uh;
            for (int j = 0; j < 3; ++j) {
                o << ", ";
                if (j == dim) o << xh;
                else if (j < P.ndim) o << coord(j);
                else o << "0.0";
            }
            o << ", c)";
            return o.str();
        };
        if (mode == TILE) {
            std::vector<double> wo, wi, wd;
            int oo, oi, od;
            // outer core must cover the core box; half-point tables must cover every tapped half point
            if (!tile_row(TO, dim, wo, oo)) { err = "nonlinear Laplacian: outer table has no core"; return false; }
            if (!TI.has_core || !TD.has_core) { err = "nonlinear Laplacian: half-point tables have no core"; return false; }
            int mlo = P.clo[dim] + oo, mhi = P.chi[dim] + oo + TO.L - 1;
            if (TI.core_lo > mlo || TI.core_hi < mhi || TD.core_lo > mlo || TD.core_hi < mhi) {
                err = "nonlinear Laplacian: half-point cores do not cover the core box";
                return false;
            }
            wi = TI.core_w; oi = TI.core_off; wd = TD.core_w; od = TD.core_off;
            code << "    " << num() << r << " = 0.0;\n";
            for (int k = 0; k < TO.L; ++k) {
                if (wo[k] == 0.0) continue;
                int m = oo + k;   // half point relative to the node
                code << "    {\n        " << num() << "uh[MOL_NVAR];\n";
                for (int v = 0; v < P.nvar; ++v) {
                    code << "        uh[" << v << "] = ";
                    bool first = true;
                    for (int q = 0; q < TI.L; ++q) {
                        if (wi[q] == 0.0) continue;
                        if (!first) code << " + ";
                        code << hexd(wi[q]) << " * " << S(v, dim, m + oi + q);
                        first = false;
                    }
                    code << ";\n";
                }
                code << "        double xh = ";
                {
                    bool first = true;
                    for (int q = 0; q < TI.L; ++q) {
                        if (wi[q] == 0.0) continue;
                        if (!first) code << " + ";
                        code << hexd(wi[q]) << " * mol_gx<" << var << "," << dim << ">(c, i" << dim << " + " << (m + oi + q) << ")";
                        first = false;
                    }
                    code << ";\n";
                }
                code << "        const " << num() << "dh = ";
                {
                    bool first = true;
                    for (int q = 0; q < TD.L; ++q) {
                        if (wd[q] == 0.0) continue;
                        if (!first) code << " + ";
                        code << hexd(wd[q]) << " * " << S(var, dim, m + od + q);
                        first = false;
                    }
                    code << ";\n";
                }
                code << "        " << r << " = fma(" << hexd(wo[k]) << ", " << fncall("uh", "xh") << " * dh, " << r << ");\n    }\n";
            }
        } else {
            const char* lin = dual ? "mol_lin_d<" : "mol_lin_g<";
            code << "    " << num() << r << " = 0.0;\n    {\n";
            code << "        const int orow = i" << dim << " - " << TO.first << ";\n";
            code << "        const int ms = __ldg(c.tabs + " << TO.soff << " + 2 * orow), mn = __ldg(c.tabs + " << TO.soff
                 << " + 2 * orow + 1);\n";
            code << "        for (int k = 0; k < mn; ++k) {\n            const int m = ms + k;\n            " << num()
                 << "uh[MOL_NVAR];\n";
            for (int v = 0; v < P.nvar; ++v)
                code << "            uh[" << v << "] = " << lin << v << "," << dim << ">" << ctx() << TI.woff << ", " << TI.soff
                     << ", " << TI.L << ", m - " << TI.first << ", i0, i1, i2);\n";
            code << "            const double xh = mol_lin_coord<" << var << "," << dim << ">(c, " << TI.woff << ", " << TI.soff << ", "
                 << TI.L << ", m - " << TI.first << ");\n";
            code << "            const " << num() << "dh = " << lin << var << "," << dim << ">" << ctx() << TD.woff << ", " << TD.soff
                 << ", " << TD.L << ", m - " << TD.first << ", i0, i1, i2);\n";
            code << "            " << r << " = fma(__ldg(c.tabw + " << TO.woff << " + (mol_i64)orow * " << TO.L << " + k), "
                 << fncall("uh", "xh") << " * dh, " << r << ");\n        }\n    }\n";
        }
        out = {r, false};
        return true;
    }

    bool run(const Rpn& rpn, std::string& result) {
        std::vector<Val> st;
        auto pop = [&](Val& v) {
            if (st.empty()) { err = "RPN stack underflow"; return false; }
            v = st.back();
            st.pop_back();
            return true;
        };
        static const std::map<std::string, std::string> unary = {
            {"sqrt", "sqrt"}, {"exp", "exp"}, {"log", "log"}, {"sin", "sin"}, {"cos", "cos"}, {"tan", "tan"},
            {"sinh", "sinh"}, {"cosh", "cosh"}, {"tanh", "tanh"}, {"abs", "fabs"}, {"asin", "asin"},
            {"acos", "acos"}, {"atan", "atan"}, {"erf", "erf"}};
        static const std::map<std::string, std::string> cmp = {{"gt", ">"}, {"ge", ">="}, {"lt", "<"},
                                                               {"le", "<="}, {"eq", "=="}, {"ne", "!="}};
        for (const std::string& tk : rpn) {
            std::vector<std::string> f = split(tk, ':');
            const std::string& op = f[0];
            Val a, b, cnd;
            if (op == "c" && f.size() == 2) {
                st.push_back({hexd(strtod(f[1].c_str(), nullptr)), false});
            } else if (op == "p" && f.size() == 2) {
                st.push_back({"c.p[" + f[1] + "]", false});
            } else if (op == "t") {
                st.push_back({"c.t", false});
            } else if (op == "x" && f.size() == 2) {
                st.push_back({fresh(coord(atoi(f[1].c_str()))), false});
            } else if (op == "u" && f.size() == 2) {
                int v = atoi(f[1].c_str());
                if (v < 0 || v >= P.nvar) { err = "bad variable in expression"; return false; }
                if (mode == GHOST) { err = "field values are not allowed in boundary data expressions"; return false; }
                if (!ucache.count(v)) {
                    if (mode == FN) ucache[v] = "uh[" + f[1] + "]";
                    else if (mode == TILE) ucache[v] = fresh(S(v, 0, 0));
                    else ucache[v] = fresh(std::string(dual ? "mol_node_d<" : "mol_node<") + f[1] + ">" + ctx() + "i0, i1, i2)");
                }
                st.push_back({ucache[v], false});
            } else if (op == "s" && f.size() == 2) {
                // a variable of t alone (it owns one node) read from another variable's equation
                int v = atoi(f[1].c_str());
                if (v < 0 || v >= P.nvar) { err = "bad variable in expression"; return false; }
                if (mode != GENERIC) { err = "point variables are read in the table-driven kernels only"; return false; }
                std::ostringstream o;
                o << (dual ? "mol_node_d<" : "mol_node<") << v << ">" << ctx();
                for (int q = 0; q < 3; ++q) o << (q < P.ndim ? P.vars[v].ilo[q] : 1) << (q < 2 ? ", " : ")");
                st.push_back({fresh(o.str()), false});
            } else if (op == "L" && f.size() == 4) {
                if (mode == FN || mode == GHOST) { err = "stencil op inside coefficient function"; return false; }
                Val o;
                if (!emit_lin(atoi(f[1].c_str()), atoi(f[2].c_str()), atoi(f[3].c_str()), o)) return false;
                st.push_back(o);
            } else if (op == "W" && f.size() == 6) {
                if (mode == FN || mode == GHOST) { err = "stencil op inside coefficient function"; return false; }
                Val o;
                if (!emit_weno(f, o)) return false;
                st.push_back(o);
            } else if (op == "M" && f.size() == 6) {
                if (mode == FN || mode == GHOST) { err = "stencil op inside coefficient function"; return false; }
                Val o;
                if (!emit_mixed(f, o)) return false;
                st.push_back(o);
            } else if (op == "N" && f.size() == 7) {
                if (mode == FN || mode == GHOST) { err = "stencil op inside coefficient function"; return false; }
                Val o;
                if (!emit_nll(f, o)) return false;
                st.push_back(o);
            } else if (op == "neg") {
                if (!pop(a)) return false;
                st.push_back({fresh("-" + asd(a)), false});
            } else if (op == "sign") {
                if (!pop(a)) return false;
                st.push_back({fresh("((" + asd(a) + " > 0.0) ? 1.0 : ((" + asd(a) + " < 0.0) ? -1.0 : 0.0))"), false});
            } else if (unary.count(op)) {
                if (!pop(a)) return false;
                st.push_back({fresh(unary.at(op) + "(" + asd(a) + ")"), false});
            } else if (op == "+" || op == "-" || op == "*" || op == "/") {
                if (!pop(b) || !pop(a)) return false;
                st.push_back({fresh(asd(a) + " " + op + " " + asd(b)), false});
            } else if (op == "pow" || op == "min" || op == "max") {
                if (!pop(b) || !pop(a)) return false;
                std::string fnm = op == "pow" ? "pow" : (op == "min" ? "fmin" : "fmax");
                st.push_back({fresh(fnm + "(" + asd(a) + ", " + asd(b) + ")"), false});
            } else if (op == "powi" && f.size() == 2) {
                if (!pop(a)) return false;
                int n = atoi(f[1].c_str());
                int m = std::abs(n);
                std::string base = asd(a), acc;
                if (m == 0) acc = "1.0";
                else {
                    // Julia lowers literal integer powers to repeated multiplication (x^2 == x*x)
                    acc = base;
                    for (int q = 1; q < m; ++q) acc = fresh(acc + " * " + base);
                }
                if (n < 0) acc = fresh("1.0 / " + acc);
                st.push_back({acc, false});
            } else if (cmp.count(op)) {
                if (!pop(b) || !pop(a)) return false;
                st.push_back({fresh(asd(a) + " " + cmp.at(op) + " " + asd(b), true), true});
            } else if (op == "and" || op == "or") {
                if (!pop(b) || !pop(a)) return false;
                st.push_back({fresh(asb(a) + (op == "and" ? " && " : " || ") + asb(b), true), true});
            } else if (op == "not") {
                if (!pop(a)) return false;
                st.push_back({fresh("!" + asb(a), true), true});
            } else if (op == "sel") {
                if (!pop(b) || !pop(a) || !pop(cnd)) return false;
                st.push_back({fresh(asb(cnd) + " ? " + asd(a) + " : " + asd(b)), false});
            } else {
                err = "unknown RPN token '" + tk + "'";
                return false;
            }
        }
        if (st.size() != 1) { err = "RPN expression does not reduce to one value"; return false; }
        result = asd(st[0]);
        return true;
    }
};

static bool uses_token(const Rpn& r, const std::string& prefix) {
    for (auto& t : r)
        if (t.compare(0, prefix.size(), prefix) == 0) return true;
    return false;
}

int generate_source(const Program& P, GenSource& G) {
    std::ostringstream pre, body;
    const int D = P.ndim, V = P.nvar;
    pre << "// ---- generated prelude ----\n";
    pre << "#define MOL_NDIM " << D << "\n#define MOL_NVAR " << V << "\n#define MOL_NPARAM " << P.nparam << "\n";
    for (int j = 0; j < 3; ++j) pre << "#define MOL_N" << j << " " << (j < D ? P.grid[j].n : 1) << "\n";
    pre << "#ifndef MOL_DIST\n#define MOL_DIST 0\n#endif\n#ifndef MOL_HALO\n#define MOL_HALO 0\n#endif\n";
    auto arr2 = [&](const char* name, auto get) {
        pre << "static __device__ constexpr int " << name << "[" << V << "][3] = {";
        for (int v = 0; v < V; ++v) {
            pre << "{";
            for (int j = 0; j < 3; ++j) pre << (j < D ? get(v, j) : 1) << (j < 2 ? "," : "");
            pre << "}" << (v + 1 < V ? "," : "");
        }
        pre << "};\n";
    };
    arr2("mol_ilo_", [&](int v, int j) { return P.vars[v].ilo[j]; });
    arr2("mol_ihi_", [&](int v, int j) { return P.vars[v].ihi[j]; });
    arr2("mol_per_", [&](int v, int j) { return P.vars[v].per[j]; });
    arr2("mol_ext_", [&](int v, int j) { return P.vars[v].ext(j); });
    pre << "static __device__ constexpr long long mol_voff_[" << V << "] = {";
    for (int v = 0; v < V; ++v) pre << P.voff[v] << (v + 1 < V ? "," : "");
    pre << "};\n";
    pre << "#define MOL_ILO(V, J) (mol_ilo_[V][J])\n#define MOL_IHI(V, J) (mol_ihi_[V][J])\n"
           "#define MOL_PER(V, J) (mol_per_[V][J])\n#define MOL_EXT(V, J) (mol_ext_[V][J])\n"
           "#define MOL_LLO0(V, c) MOL_ILO(V, 0)\n"
           "#if MOL_DIST\n"
           "#define MOL_VOFF(V, c) ((long long)(V) * (c).vstride)\n#define MOL_LLO(V, c) ((c).loc_lo)\n"
           "#else\n"
           "#define MOL_VOFF(V, c) (mol_voff_[V])\n#define MOL_LLO(V, c) MOL_ILO(V, MOL_NDIM - 1)\n"
           "#endif\n";
    {   // doubles per variable in one plane normal to the last (split) dimension
        long long pmax = 1;
        pre << "static __device__ constexpr long long mol_plane_[" << V << "] = {";
        for (int v = 0; v < V; ++v) {
            long long pl = 1;
            for (int j = 0; j + 1 < D; ++j) pl *= P.vars[v].ext(j);
            pmax = std::max(pmax, pl);
            pre << pl << (v + 1 < V ? "," : "");
        }
        pre << "};\n#define MOL_PLANE(V) (mol_plane_[V])\n#define MOL_PLANE_MAX " << pmax << "LL\n";
    }
    for (int j = 0; j < 3; ++j) {
        int lomax = 1, himin = 1;
        if (j < D) {
            lomax = P.vars[0].ilo[j];
            himin = P.vars[0].ihi[j];
            for (int v = 1; v < V; ++v) {
                lomax = std::max(lomax, P.vars[v].ilo[j]);
                himin = std::min(himin, P.vars[v].ihi[j]);
            }
        }
        pre << "#define MOL_ILO_MAX" << j << " " << lomax << "\n#define MOL_IHI_MIN" << j << " " << himin << "\n";
    }

    // ---- ghost rules ------------------------------------------------------------------------
    body << "// ---- generated ghost rules (generate_bc_eqs.jl:313-328, 336-392) ----\n";
    for (int v = 0; v < V; ++v)
        for (int j = 0; j < 3; ++j)
            body << "template <> __device__ double mol_ghost<" << v << "," << j
                 << ">(const MolIn& in, const MolCtx& c, int i0, int i1, int i2);\n";
    for (int v = 0; v < V; ++v) {
        for (int j = 0; j < 3; ++j) {
            body << "template <> __device__ double mol_ghost<" << v << "," << j
                 << ">(const MolIn& in, const MolCtx& c, int i0, int i1, int i2) {\n";
            bool any = false;
            for (const Ghost& g : P.ghosts) {
                if (g.var != v || g.dim != j) continue;
                if (!any) body << "    switch (i" << j << ") {\n";
                any = true;
                body << "    case " << g.node << ": {\n";
                int reach[3] = {0, 0, 0};
                Emitter E(P, GHOST, reach);
                std::string res;
                if (!E.run(g.expr, res)) return fail(MOL_E_PARSE, "ghost rule: " + E.err);
                body << E.code.str();
                body << "    double r = " << res << ";\n";
                for (size_t kt = 0; kt < g.taps.size(); ++kt) {
                    const GhostTap& tp = g.taps[kt];
                    std::string coef = hexd(tp.coef);
                    if (!g.tapexpr.empty()) {            // coefficient expression (parameters, t, boundary coordinates)
                        body << "    {\n";
                        Emitter EC(P, GHOST, reach);
                        if (!EC.run(g.tapexpr[kt], coef)) return fail(MOL_E_PARSE, "ghost tap coefficient: " + EC.err);
                        body << EC.code.str() << "    const double ck" << kt << " = " << coef << ";\n";
                        coef = "ck" + std::to_string(kt);
                    }
                    body << "    r = fma(" << coef << ", mol_node<" << tp.var << ">(in, c, ";
                    for (int q = 0; q < 3; ++q) {
                        if (q == j) body << tp.node; else body << "i" << q;
                        body << (q < 2 ? ", " : "");
                    }
                    body << "), r);\n";
                    if (!g.tapexpr.empty()) body << "    }\n";
                }
                body << "    return r; }\n";
            }
            if (any) body << "    default: break;\n    }\n";
            body << "    return 0.0;\n}\n";
        }
    }

    // ---- coefficient functions -----------------------------------------------------------------
    for (auto& kv : P.fns) {
        body << "__device__ __forceinline__ double mol_fn_" << kv.first
             << "(const double* uh, double xh0, double xh1, double xh2, const MolCtx& c) {\n";
        int reach[3] = {0, 0, 0};
        Emitter E(P, FN, reach);
        std::string res;
        if (!E.run(kv.second, res)) return fail(MOL_E_PARSE, "coefficient function: " + E.err);
        body << E.code.str() << "    return " << res << ";\n}\n";
    }

    // ---- generic equations ------------------------------------------------------------------------
    body << "template <int V> __device__ __forceinline__ double mol_eq_generic(const MolIn& in, const MolCtx& c, int i0, int i1, int i2);\n";
    for (int v = 0; v < V; ++v) {
        body << "template <> __device__ __forceinline__ double mol_eq_generic<" << v
             << ">(const MolIn& in, const MolCtx& c, int i0, int i1, int i2) {\n";
        int reach[3] = {0, 0, 0};
        Emitter E(P, GENERIC, reach);
        std::string res;
        if (!E.run(P.eqs[v], res)) return fail(MOL_E_PARSE, "equation " + std::to_string(v) + ": " + E.err);
        body << E.code.str() << "    return " << res << ";\n}\n";
    }

    // ---- Jacobian-vector product: the same ghost rules, coefficient functions and equations on dual numbers
    //      (kernels/mol_jvp.cuh), table-driven forms only ------------------------------------------------------------
    body << "#if MOL_KERNEL_JVP\n";
    for (int v = 0; v < V; ++v)
        for (int j = 0; j < 3; ++j)
            body << "template <> __device__ MolDual mol_ghost_d<" << v << "," << j
                 << ">(const MolIn& in, const MolJv& jv, const MolCtx& c, int i0, int i1, int i2);\n";
    for (int v = 0; v < V; ++v) {
        for (int j = 0; j < 3; ++j) {
            body << "template <> __device__ MolDual mol_ghost_d<" << v << "," << j
                 << ">(const MolIn& in, const MolJv& jv, const MolCtx& c, int i0, int i1, int i2) {\n";
            bool any = false;
            for (const Ghost& g : P.ghosts) {
                if (g.var != v || g.dim != j) continue;
                if (!any) body << "    switch (i" << j << ") {\n";
                any = true;
                body << "    case " << g.node << ": {\n";
                int reach[3] = {0, 0, 0};
                Emitter E(P, GHOST, reach);                  // boundary data: no field values, plain doubles
                std::string res;
                if (!E.run(g.expr, res)) return fail(MOL_E_PARSE, "ghost rule: " + E.err);
                body << E.code.str();
                body << "    MolDual r = " << res << ";\n";
                for (size_t kt = 0; kt < g.taps.size(); ++kt) {
                    const GhostTap& tp = g.taps[kt];
                    std::string coef = hexd(tp.coef);
                    if (!g.tapexpr.empty()) {
                        body << "    {\n";
                        Emitter EC(P, GHOST, reach);
                        if (!EC.run(g.tapexpr[kt], coef)) return fail(MOL_E_PARSE, "ghost tap coefficient: " + EC.err);
                        body << EC.code.str() << "    const double ck" << kt << " = " << coef << ";\n";
                        coef = "ck" + std::to_string(kt);
                    }
                    body << "    r = fma(" << coef << ", mol_node_d<" << tp.var << ">(in, jv, c, ";
                    for (int q = 0; q < 3; ++q) {
                        if (q == j) body << tp.node; else body << "i" << q;
                        body << (q < 2 ? ", " : "");
                    }
                    body << "), r);\n";
                    if (!g.tapexpr.empty()) body << "    }\n";
                }
                body << "    return r; }\n";
            }
            if (any) body << "    default: break;\n    }\n";
            body << "    return MolDual(0.0);\n}\n";
        }
    }
    for (auto& kv : P.fns) {
        body << "__device__ __forceinline__ MolDual mol_fnd_" << kv.first
             << "(const MolDual* uh, double xh0, double xh1, double xh2, const MolCtx& c) {\n";
        int reach[3] = {0, 0, 0};
        Emitter E(P, FN, reach, true);
        std::string res;
        if (!E.run(kv.second, res)) return fail(MOL_E_PARSE, "coefficient function: " + E.err);
        body << E.code.str() << "    return " << res << ";\n}\n";
    }
    body << "template <int V> __device__ __forceinline__ MolDual mol_eq_jvp(const MolIn& in, const MolJv& jv, const MolCtx& c, "
            "int i0, int i1, int i2);\n";
    for (int v = 0; v < V; ++v) {
        body << "template <> __device__ __forceinline__ MolDual mol_eq_jvp<" << v
             << ">(const MolIn& in, const MolJv& jv, const MolCtx& c, int i0, int i1, int i2) {\n";
        int reach[3] = {0, 0, 0};
        Emitter E(P, GENERIC, reach, true);
        std::string res;
        if (!E.run(P.eqs[v], res)) return fail(MOL_E_PARSE, "equation " + std::to_string(v) + " (JVP): " + E.err);
        body << E.code.str() << "    return " << res << ";\n}\n";
    }
    body << "#endif  // MOL_KERNEL_JVP\n";

    // ---- tiled equations (literal weights) ------------------------------------------------------
    TileCfg& T = G.tile;
    T.enabled = false;
    std::ostringstream tbody;
    if (P.has_core) {
        bool ok = true;
        // all variables must share the interior box so one tile serves every equation
        for (int v = 1; v < V && ok; ++v)
            for (int j = 0; j < D; ++j)
                if (P.vars[v].ilo[j] != P.vars[0].ilo[j] || P.vars[v].ihi[j] != P.vars[0].ihi[j]) ok = false;
        int reach[3] = {0, 0, 0};
        tbody << "template <int V> __device__ __forceinline__ double mol_eq_tile(const double* __restrict__ sm, "
                 "const double* __restrict__ wsm, const MolCtx& c, "
                 "int lx, int ly, int lz, int i0, int i1, int i2, double xc0, double xc1, double xc2);\n";
        bool use_x[3] = {false, false, false};
        for (int v = 0; v < V && ok; ++v) {
            Emitter E(P, TILE, reach);
            std::string res;
            if (!E.run(P.eqs[v], res)) { ok = false; break; }
            tbody << "template <> __device__ __forceinline__ double mol_eq_tile<" << v
                  << ">(const double* __restrict__ sm, const double* __restrict__ wsm, const MolCtx& c, int lx, int ly, int lz, "
                     "int i0, int i1, int i2, double xc0, double xc1, double xc2) {\n"
                  << E.code.str() << "    return " << res << ";\n}\n";
            for (int j = 0; j < 3; ++j) use_x[j] = use_x[j] || E.uses_coord[j];
        }
        for (int j = 0; j < 3; ++j) pre << "#define MOL_USE_X" << j << " " << (use_x[j] ? 1 : 0) << "\n";
        // the same equations on dual numbers for the tiled Jacobian-vector product (u tile + v tile, kernels/mol_tiled.cuh
        // MOL_KERNEL_JVP); programs whose tiled form has no dual twin keep the table-driven J*v
        T.jvp = false;
        if (ok && D <= 2) {
            std::ostringstream jb;
            bool jok = true;
            jb << "#if MOL_KERNEL_JVP\ntemplate <int V> __device__ __forceinline__ MolDual mol_eq_tile_d(const double* __restrict__ sm, "
                  "const double* __restrict__ smv, const double* __restrict__ wsm, const MolCtx& c, "
                  "int lx, int ly, int lz, int i0, int i1, int i2, double xc0, double xc1, double xc2);\n";
            int reach_d[3] = {0, 0, 0};
            for (int v = 0; v < V && jok; ++v) {
                Emitter E(P, TILE, reach_d, true);
                std::string res;
                if (!E.run(P.eqs[v], res)) { jok = false; break; }
                jb << "template <> __device__ __forceinline__ MolDual mol_eq_tile_d<" << v
                   << ">(const double* __restrict__ sm, const double* __restrict__ smv, const double* __restrict__ wsm, const MolCtx& c, "
                      "int lx, int ly, int lz, int i0, int i1, int i2, double xc0, double xc1, double xc2) {\n"
                   << E.code.str() << "    return " << res << ";\n}\n";
            }
            jb << "#endif  // MOL_KERNEL_JVP\n";
            if (jok) { tbody << jb.str(); T.jvp = true; }
        }
        if (ok) {
            T.enabled = true;
            for (int j = 0; j < 3; ++j) T.r[j] = (j < D) ? reach[j] : 0;
            T.r0p = (T.r[0] + 1) / 2 * 2;
            T.vx = 2;
            T.nthreads = 256;
            T.stages = 2;
            auto env_flag = [](const char* name, bool dflt) {
                const char* e = getenv(name);
                return (e && *e) ? atoi(e) != 0 : dflt;
            };
            if (D == 1) { T.tx = 2048; T.ty = 1; T.tz = 1; }
            // 2-D: measured on B200 at 4096^2 x 2 species (profiles/r01_tile_sweep.md): 64 x 16 tiles, 3 TMA
            // stages, register cap for 4 CTAs/SM -> 91.5 us = 89 % of the measured HBM copy rate
            else if (D == 2) { T.tx = 64; T.ty = 16; T.tz = 1; T.stages = 3; T.min_ctas = 4; }
            // 3-D: xy tiles marching along z (profiles/r01_3d.md): 128 x 8 tiles, chunks of 8 planes, a ring of
            // 2*r + 3 planes (two planes of TMA prefetch), register cap for 4 CTAs/SM.  MOL_TILE_ZMARCH=0 selects the
            // brick kernel (64 x 8 x 4 tiles) instead.
            else {
                T.zmarch = env_flag("MOL_TILE_ZMARCH", true);
                if (T.zmarch) { T.tx = 128; T.ty = 8; T.tz = 8; T.ring = 2 * T.r[2] + 3; T.min_ctas = 4; }
                else { T.tx = 64; T.ty = 8; T.tz = 4; }
            }
            // 2-D programs with staged per-node weight records (non-uniform axes): taller tiles amortise the column
            // records and the two barriers per tile over twice the rows (4 rows per thread), and the arithmetic wants
            // ~128 registers.  Measured on a B200, 4097^2 non-uniform upwind Burgers (config 3), 256 threads:
            // 64x16 / 3 stages / 3 CTAs per SM 219.6 us; 64x32 / 2 stages / 2 CTAs 159.7 us; 128x16 / 2 / 2 166.9 us;
            // 64x32 with 512 threads (64-register cap) 321.8 us.
            if (D == 2 && (P.wrec_stride[0] > 0 || P.wrec_stride[1] > 0)) { T.ty = 32; T.stages = 2; T.min_ctas = 2; }
            // tuning overrides (experiments only; the defaults above are the shipped configuration)
            auto env_int = [](const char* name, int dflt) {
                const char* e = getenv(name);
                return (e && *e) ? atoi(e) : dflt;
            };
            T.tx = env_int("MOL_TILE_TX", T.tx);
            if (D >= 2) T.ty = env_int("MOL_TILE_TY", T.ty);
            if (D >= 3) T.tz = env_int("MOL_TILE_TZ", T.tz);
            T.stages = env_int("MOL_TILE_STAGES", T.stages);
            T.nthreads = env_int("MOL_TILE_THREADS", T.nthreads);
            T.min_ctas = env_int("MOL_TILE_MINCTAS", T.min_ctas);
            if (T.zmarch) {
                T.ring = env_int("MOL_TILE_RING", T.ring);
                T.l2_ahead = env_int("MOL_TILE_L2AHEAD", T.l2_ahead);
                if (T.ring < 2 * T.r[2] + 2 || T.ring > 32) return fail(MOL_E_ARG, "z-march ring must hold at least 2*r + 2 planes");
            }
            {   // thread layout must cover the tile exactly: rows per thread = TY / (NTHREADS / min(TX/VX, NTHREADS))
                const int ntx = T.tx / T.vx, ntxt = std::min(ntx, T.nthreads);
                if (T.stages < 2 || T.tx % T.vx || ntx % ntxt || T.nthreads % ntxt || (D >= 2 && T.ty % (T.nthreads / ntxt)))
                    return fail(MOL_E_ARG, "tile configuration does not divide evenly among the CTA's threads");
            }
            const bool wrec = P.wrec_stride[0] > 0 || P.wrec_stride[1] > 0;
            if (wrec && D == 1 && !getenv("MOL_TILE_TX")) {
                // 1-D with staged records: (8 nvar + 8 stride) bytes per node and stage; keep a CTA near 48 KB (4 CTAs/SM)
                const int per_node = 8 * V + 8 * P.wrec_stride[0];
                int tx = 2048;
                while (tx > 256 && (size_t)T.stages * tx * per_node > 56 * 1024) tx /= 2;
                T.tx = tx;
            }
            bool align = true;
            for (int v = 0; v < V; ++v)
                if (P.vars[v].ext(0) % 2 != 0 || P.voff[v] % 2 != 0) align = false;
            T.vec_store = align && ((P.clo[0] - P.vars[0].ilo[0]) % 2 == 0);
            T.tma = align && D >= 2 && ((P.clo[0] - T.r0p - P.vars[0].ilo[0]) % 2 == 0 || true);
            // programs with staged per-node records run the cp.async pipeline (the records travel in its commit groups)
            if (wrec && !T.zmarch && !getenv("MOL_TILE_FORCE_TMA")) T.tma = false;
            T.wstage_doubles = (wrec && !T.zmarch && D <= 2)
                                   ? (size_t)P.wrec_stride[0] * ((T.tx + P.wrec_hl[0] + P.wrec_hh[0] + 1) / 2 * 2) +
                                         (D >= 2 ? (size_t)P.wrec_stride[1] * (T.ty + P.wrec_hl[1] + P.wrec_hh[1]) : 0)
                                   : 0;
            size_t cells = (size_t)(T.tx + 2 * T.r0p) * (D >= 2 ? T.ty + 2 * T.r[1] : 1) *
                           (D >= 3 && !T.zmarch ? T.tz + 2 * T.r[2] : 1);
            T.tile_stride_doubles = (cells * 8 + 127) / 128 * 128 / 8;
        }
    }
    pre << "#define MOL_HAVE_TILE " << (T.enabled ? 1 : 0) << "\n";
    if (T.enabled) {
        pre << "#define MOL_TX " << T.tx << "\n#define MOL_TY " << T.ty << "\n#define MOL_TZ " << T.tz << "\n"
            << "#define MOL_VX " << T.vx << "\n#define MOL_NTHREADS " << T.nthreads << "\n#define MOL_STAGES " << T.stages
            << "\n#define MOL_R0 " << T.r[0] << "\n#define MOL_R1 " << T.r[1] << "\n#define MOL_R2 " << T.r[2]
            << "\n#define MOL_R0P " << T.r0p << "\n#define MOL_VEC_ST " << (T.vec_store ? 1 : 0) << "\n#define MOL_ZMARCH "
            << (T.zmarch ? 1 : 0) << "\n#define MOL_RING " << T.ring << "\n#define MOL_L2_AHEAD " << T.l2_ahead << "\n";
        if (P.wrec_stride[0] > 0 || P.wrec_stride[1] > 0)
            for (int j = 0; j < 2; ++j)
                pre << "#define MOL_WRS" << j << " " << P.wrec_stride[j] << "\n#define MOL_WHL" << j << " " << P.wrec_hl[j]
                    << "\n#define MOL_WHH" << j << " " << P.wrec_hh[j] << "\n#define MOL_WOFF" << j << " " << P.wrec_off[j]
                    << "\n#define MOL_WLO" << j << " " << P.wrec_lo[j] << "\n#define MOL_WNREC" << j << " " << P.wrec_n[j] << "\n";
        for (int j = 0; j < 3; ++j)
            pre << "#define MOL_CLO" << j << " " << (j < D ? P.clo[j] : 1) << "\n#define MOL_CHI" << j << " "
                << (j < D ? P.chi[j] : 1) << "\n";
        body << "#if MOL_KERNEL_TILED\n" << tbody.str() << "#endif\n";
    }
    G.prelude = pre.str();
    G.body = body.str();
    return MOL_OK;
}

}  // namespace mol
