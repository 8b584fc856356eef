// fpsb_symbolic.cpp — host-side symbolic analysis for the LDLt path.
//
// Replaces `ldl_analyze(Symmetric(sparse(rows, cols, vals, N, N), :U))` at
// /root/reference/src/solve_two_systems_struct.jl:343-348 (LDLFactorizations.jl):
//   * fill-reducing ordering: an approximate-minimum-degree ordering on the pattern of K + K'
//     (quotient graph, approximate external degrees, element absorption, mass elimination,
//     supervariable detection, dense-row deferral, assembly-tree postorder — the algorithm of
//     Amestoy/Davis/Duff that LDLFactorizations reaches through AMD.jl).
//     Attribution: amd_order below follows the structure and variable naming (Pe, Len, Nv, Elen,
//     Degree, W, Iw, pfree, wflg, FLIP, clear_flag, mindeg, nel, lemax) of SuiteSparse AMD's
//     amd_2 routine — AMD, Copyright (c) 1996-2022, Timothy A. Davis, Patrick R. Amestoy and
//     Iain S. Duff, BSD-3-Clause licence — re-typed here from the algorithm as published in
//     "An approximate minimum degree ordering algorithm", SIAM J. Matrix Anal. Appl. 17(4), 1996 and
//     "Algorithm 837: AMD", ACM TOMS 30(3), 2004.  SuiteSparse is not vendored (nor present in
//     /root/reference), so the result is NOT verified bit-identical to libamd, which is why
//     fpsb_ldlt_analyze also accepts an explicit P like `ldl_analyze(A, P)`.
//   * elimination tree (Liu, path compression) + exact column structures of L by bottom-up
//     child merging.  Given P, (parent, Lnz, Lp, Li) are canonical and are checked bit-exactly
//     against the oracle's row-subtree walk (a different algorithm).
//   * B200 plan: relaxed supernodes (dense panels), level-ordered task list, left-looking update
//     pairs with precomputed relative indices, and the COO -> panel assembly map that replaces the
//     reference's per-refactor `sparse(rows, cols, vals)` (src/solve_linear_system.jl:233).
#include "fpsb_symbolic.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <stdexcept>

namespace fpsb {

namespace {
constexpr int EMPTY = -1;
inline int64_t FLIP(int64_t i) { return -i - 2; }

int clear_flag(int wflg, int wbig, std::vector<int> &W, int n) {
    if (wflg < 2 || wflg >= wbig) {
        for (int x = 0; x < n; x++) if (W[x] != 0) W[x] = 1;
        wflg = 2;
    }
    return wflg;
}
}  // namespace

void amd_order(int n, const std::vector<int64_t> &Ap, const std::vector<int> &Ai, std::vector<int> &Pout) {
    Pout.assign((size_t)n, 0);
    if (n == 0) return;
    const int64_t nz = Ap[(size_t)n];
    int64_t iwlen = nz + nz / 5 + n + 16;
    std::vector<int> Iw((size_t)iwlen);
    std::vector<int64_t> Pe((size_t)n);
    std::vector<int> Len((size_t)n), Nv((size_t)n), Next((size_t)n), Last((size_t)n), Head((size_t)n),
        Elen((size_t)n), Degree((size_t)n), W((size_t)n);
    int64_t pfree = 0;
    for (int j = 0; j < n; j++) {
        Pe[j] = pfree;
        Len[j] = (int)(Ap[j + 1] - Ap[j]);
        for (int64_t p = Ap[j]; p < Ap[j + 1]; p++) Iw[(size_t)pfree++] = Ai[(size_t)p];
    }
    const int wbig = INT32_MAX - n;
    for (int i = 0; i < n; i++) {
        Last[i] = EMPTY; Head[i] = EMPTY; Next[i] = EMPTY; Nv[i] = 1; W[i] = 1; Elen[i] = 0;
        Degree[i] = Len[i];
    }
    int wflg = clear_flag(0, wbig, W, n);
    int mindeg = 0, nel = 0, lemax = 0;
    int dense = (int)(10.0 * std::sqrt((double)n));
    dense = std::max(16, dense);
    dense = std::min(n, dense);
    for (int i = 0; i < n; i++) {
        int deg = Degree[i];
        if (deg == 0) {
            Elen[i] = (int)FLIP(1); nel++; Pe[i] = EMPTY; W[i] = 0;
        } else if (deg > dense) {
            Nv[i] = 0; Elen[i] = EMPTY; nel++; Pe[i] = EMPTY;
        } else {
            int inext = Head[deg];
            if (inext != EMPTY) Last[inext] = i;
            Next[i] = inext;
            Head[deg] = i;
        }
    }
    auto remove_from_list = [&](int i) {
        int ilast = Last[i], inext = Next[i];
        if (inext != EMPTY) Last[inext] = ilast;
        if (ilast != EMPTY) Next[ilast] = inext;
        else Head[Degree[i]] = inext;
    };
    while (nel < n) {
        int deg, me = EMPTY;
        for (deg = mindeg; deg < n; deg++) { me = Head[deg]; if (me != EMPTY) break; }
        if (me == EMPTY) throw std::runtime_error("amd_order: internal error (no pivot)");
        mindeg = deg;
        int inext = Next[me];
        if (inext != EMPTY) Last[inext] = EMPTY;
        Head[deg] = inext;
        const int elenme = Elen[me];
        int nvpiv = Nv[me];
        nel += nvpiv;
        Nv[me] = -nvpiv;
        int degme = 0;
        int64_t pme1, pme2;
        if (elenme == 0) {
            pme1 = Pe[me]; pme2 = pme1 - 1;
            for (int64_t p = pme1; p <= pme1 + Len[me] - 1; p++) {
                int i = Iw[(size_t)p];
                int nvi = Nv[i];
                if (nvi > 0) {
                    degme += nvi; Nv[i] = -nvi; Iw[(size_t)(++pme2)] = i;
                    remove_from_list(i);
                }
            }
        } else {
            int64_t p = Pe[me];
            pme1 = pfree;
            const int slenme = Len[me] - elenme;
            for (int knt1 = 1; knt1 <= elenme + 1; knt1++) {
                int e, ln;
                int64_t pj;
                if (knt1 > elenme) { e = me; pj = p; ln = slenme; }
                else { e = Iw[(size_t)p++]; pj = Pe[e]; ln = Len[e]; }
                for (int knt2 = 1; knt2 <= ln; knt2++) {
                    int i = Iw[(size_t)pj++];
                    int nvi = Nv[i];
                    if (nvi > 0) {
                        if (pfree >= iwlen) {
                            // garbage collection
                            Pe[me] = p; Len[me] -= knt1; if (Len[me] == 0) Pe[me] = EMPTY;
                            Pe[e] = pj; Len[e] = ln - knt2; if (Len[e] == 0) Pe[e] = EMPTY;
                            for (int j = 0; j < n; j++) {
                                int64_t pn = Pe[j];
                                if (pn >= 0) { Pe[j] = Iw[(size_t)pn]; Iw[(size_t)pn] = (int)FLIP(j); }
                            }
                            int64_t psrc = 0, pdst = 0, pend = pme1 - 1;
                            while (psrc <= pend) {
                                int j = (int)FLIP(Iw[(size_t)psrc++]);
                                if (j >= 0) {
                                    Iw[(size_t)pdst] = (int)Pe[j];
                                    Pe[j] = pdst++;
                                    int lenj = Len[j];
                                    for (int k3 = 0; k3 <= lenj - 2; k3++) Iw[(size_t)pdst++] = Iw[(size_t)psrc++];
                                }
                            }
                            int64_t p1 = pdst;
                            for (psrc = pme1; psrc <= pfree - 1; psrc++) Iw[(size_t)pdst++] = Iw[(size_t)psrc];
                            pme1 = p1; pfree = pdst; pj = Pe[e]; p = Pe[me];
                            if (pfree >= iwlen) {   // still no room: grow the workspace
                                iwlen = iwlen + iwlen / 2 + n;
                                Iw.resize((size_t)iwlen);
                            }
                        }
                        degme += nvi; Nv[i] = -nvi; Iw[(size_t)pfree++] = i;
                        remove_from_list(i);
                    }
                }
                if (e != me) { Pe[e] = FLIP(me); W[e] = 0; }
            }
            pme2 = pfree - 1;
        }
        Degree[me] = degme; Pe[me] = pme1; Len[me] = (int)(pme2 - pme1 + 1);
        Elen[me] = (int)FLIP(nvpiv + degme);
        wflg = clear_flag(wflg, wbig, W, n);
        for (int64_t pme = pme1; pme <= pme2; pme++) {
            int i = Iw[(size_t)pme];
            int eln = Elen[i];
            if (eln > 0) {
                int nvi = -Nv[i];
                int wnvi = wflg - nvi;
                for (int64_t p = Pe[i]; p <= Pe[i] + eln - 1; p++) {
                    int e = Iw[(size_t)p];
                    int we = W[e];
                    if (we >= wflg) we -= nvi;
                    else if (we != 0) we = Degree[e] + wnvi;
                    W[e] = we;
                }
            }
        }
        for (int64_t pme = pme1; pme <= pme2; pme++) {
            int i = Iw[(size_t)pme];
            int64_t p1 = Pe[i], p2 = p1 + Elen[i] - 1, pn = p1;
            unsigned hash = 0;
            int d = 0;
            for (int64_t p = p1; p <= p2; p++) {
                int e = Iw[(size_t)p];
                int we = W[e];
                if (we != 0) {
                    int dext = we - wflg;
                    if (dext > 0) { d += dext; Iw[(size_t)pn++] = e; hash += (unsigned)e; }
                    else { Pe[e] = FLIP(me); W[e] = 0; }   // aggressive absorption
                }
            }
            Elen[i] = (int)(pn - p1 + 1);
            int64_t p3 = pn, p4 = p1 + Len[i];
            for (int64_t p = p2 + 1; p < p4; p++) {
                int j = Iw[(size_t)p];
                int nvj = Nv[j];
                if (nvj > 0) { d += nvj; Iw[(size_t)pn++] = j; hash += (unsigned)j; }
            }
            if (Elen[i] == 1 && p3 == pn) {
                Pe[i] = FLIP(me);
                int nvi = -Nv[i];
                degme -= nvi; nvpiv += nvi; nel += nvi; Nv[i] = 0; Elen[i] = EMPTY;
            } else {
                Degree[i] = std::min(Degree[i], d);
                Iw[(size_t)pn] = Iw[(size_t)p3];
                Iw[(size_t)p3] = Iw[(size_t)p1];
                Iw[(size_t)p1] = me;
                Len[i] = (int)(pn - p1 + 1);
                int hsh = (int)(hash % (unsigned)n);
                int j = Head[hsh];
                if (j <= EMPTY) { Next[i] = (int)FLIP(j); Head[hsh] = (int)FLIP(i); }
                else { Next[i] = Last[j]; Last[j] = i; }
                Last[i] = hsh;
            }
        }
        Degree[me] = degme;
        lemax = std::max(lemax, degme);
        wflg += lemax;
        wflg = clear_flag(wflg, wbig, W, n);
        for (int64_t pme = pme1; pme <= pme2; pme++) {
            int i = Iw[(size_t)pme];
            if (Nv[i] < 0) {
                int hsh = Last[i];
                int j = Head[hsh];
                if (j == EMPTY) { i = EMPTY; }
                else if (j < EMPTY) { i = (int)FLIP(j); Head[hsh] = EMPTY; }
                else { i = Last[j]; Last[j] = EMPTY; }
                while (i != EMPTY && Next[i] != EMPTY) {
                    int ln = Len[i], eln = Elen[i];
                    for (int64_t p = Pe[i] + 1; p <= Pe[i] + ln - 1; p++) W[Iw[(size_t)p]] = wflg;
                    int jlast = i;
                    j = Next[i];
                    while (j != EMPTY) {
                        bool ok = (Len[j] == ln) && (Elen[j] == eln);
                        for (int64_t p = Pe[j] + 1; ok && p <= Pe[j] + ln - 1; p++)
                            if (W[Iw[(size_t)p]] != wflg) ok = false;
                        if (ok) {
                            Pe[j] = FLIP(i); Nv[i] += Nv[j]; Nv[j] = 0; Elen[j] = EMPTY;
                            j = Next[j]; Next[jlast] = j;
                        } else { jlast = j; j = Next[j]; }
                    }
                    wflg++;
                    i = Next[i];
                }
            }
        }
        int64_t p = pme1;
        int nleft = n - nel;
        for (int64_t pme = pme1; pme <= pme2; pme++) {
            int i = Iw[(size_t)pme];
            int nvi = -Nv[i];
            if (nvi > 0) {
                Nv[i] = nvi;
                int d = Degree[i] + degme - nvi;
                d = std::min(d, nleft - nvi);
                d = std::max(d, 0);
                int inx = Head[d];
                if (inx != EMPTY) Last[inx] = i;
                Next[i] = inx; Last[i] = EMPTY; Head[d] = i;
                mindeg = std::min(mindeg, d);
                Degree[i] = d;
                Iw[(size_t)p++] = i;
            }
        }
        Nv[me] = nvpiv;
        Len[me] = (int)(p - pme1);
        if (Len[me] == 0) { Pe[me] = EMPTY; W[me] = 0; }
        if (elenme != 0) pfree = p;
    }

    // ---- post-processing: assembly tree, postorder, permutation -------------------------------
    std::vector<int> par((size_t)n);
    for (int i = 0; i < n; i++) par[i] = (int)FLIP(Pe[i]);
    for (int i = 0; i < n; i++) Elen[i] = (int)FLIP(Elen[i]);
    for (int i = 0; i < n; i++) {
        if (Nv[i] == 0) {
            int j = par[i];
            if (j == EMPTY) continue;
            while (Nv[j] == 0) j = par[j];
            int e = j;
            j = i;
            while (Nv[j] == 0) { int jn = par[j]; par[j] = e; j = jn; }
        }
    }
    // children lists of elements (ascending index), largest child (by front size) moved last
    std::vector<int> child((size_t)n, EMPTY), sib((size_t)n, EMPTY), order((size_t)n, EMPTY);
    for (int j = n - 1; j >= 0; j--) {
        if (Nv[j] > 0) {
            int pa = par[j];
            if (pa != EMPTY) { sib[j] = child[pa]; child[pa] = j; }
        }
    }
    for (int i = 0; i < n; i++) {
        if (Nv[i] > 0 && child[i] != EMPTY) {
            int fprev = EMPTY, maxfr = EMPTY, bigfprev = EMPTY, bigf = EMPTY;
            for (int f = child[i]; f != EMPTY; f = sib[f]) {
                int fr = Elen[f];
                if (fr >= maxfr) { maxfr = fr; bigfprev = fprev; bigf = f; }
                fprev = f;
            }
            int fnext = sib[bigf];
            if (fnext != EMPTY) {
                if (bigfprev == EMPTY) child[i] = fnext;
                else sib[bigfprev] = fnext;
                sib[bigf] = EMPTY;
                sib[fprev] = bigf;
            }
        }
    }
    {
        std::vector<int> stack((size_t)n);
        int k = 0;
        for (int i = 0; i < n; i++) {
            if (par[i] == EMPTY && Nv[i] > 0) {
                int head = 0;
                stack[0] = i;
                while (head >= 0) {
                    int v = stack[(size_t)head];
                    if (child[v] != EMPTY) {
                        // push children so that the first child is on top
                        int cnt = 0;
                        for (int f = child[v]; f != EMPTY; f = sib[f]) cnt++;
                        int h = head + cnt;
                        for (int f = child[v]; f != EMPTY; f = sib[f]) stack[(size_t)(h--)] = f;
                        head += cnt;
                        child[v] = EMPTY;
                    } else {
                        head--;
                        order[v] = k++;
                    }
                }
            }
        }
    }
    std::vector<int> head2((size_t)n, EMPTY), nxt((size_t)n, EMPTY);
    for (int e = 0; e < n; e++) { int k = order[e]; if (k != EMPTY) head2[k] = e; }
    int cnt = 0;
    for (int k = 0; k < n; k++) {
        int e = head2[k];
        if (e == EMPTY) break;
        nxt[e] = cnt;
        cnt += Nv[e];
    }
    for (int i = 0; i < n; i++) {
        if (Nv[i] == 0) {
            int e = par[i];
            if (e != EMPTY) { nxt[i] = nxt[e]; nxt[e]++; }
            else nxt[i] = cnt++;
        }
    }
    std::vector<char> seen((size_t)n, 0);
    for (int i = 0; i < n; i++) {
        int k = nxt[i];
        if (k < 0 || k >= n || seen[(size_t)k]) throw std::runtime_error("amd_order: invalid permutation");
        seen[(size_t)k] = 1;
        Pout[(size_t)k] = i;
    }
}

// ------------------------------------------------------------------------------------------------
// One-way dissection by BFS level sets + cyclic-reduction order of the separators.
// A B200-oriented alternative to the minimum-degree default: AMD minimises fill but, on band-like
// KKT structure, yields an elimination tree that is one long chain (14 000+ dependent supernodes at
// n = 1M), which serialises the GPU factorisation.  Here the graph is cut into `nparts` pieces by
// single BFS levels (each level is a vertex separator), piece interiors are ordered independently
// (minimum degree on the induced subgraph) and the separators are ordered like cyclic reduction
// (odd ones first, then those = 2 mod 4, ...), so the dependency depth is
// O(depth of one piece + log2(nparts)) instead of O(N / supernode width).
// Passed to the analysis as an explicit P (the `ldl_analyze(A, P)` form).
// ------------------------------------------------------------------------------------------------
void dissection_order(int n, const std::vector<int64_t> &Ap, const std::vector<int> &Ai, int nparts_target,
                      std::vector<int> &Pout) {
    Pout.clear();
    Pout.reserve((size_t)n);
    if (n == 0) return;
    if (nparts_target <= 0) nparts_target = std::max(1, std::min(1024, n / 6000));
    const int part_size = std::max(64, n / std::max(1, nparts_target));
    std::vector<int> level((size_t)n, -1), comp_nodes, queue((size_t)n);
    std::vector<char> visited((size_t)n, 0);
    std::vector<int> local((size_t)n, -1);
    std::vector<std::vector<int>> separators_all;   // separators of all components, in CR order at the end
    std::vector<std::pair<int, int>> sep_rank;       // (trailing zeros, running id)
    auto bfs = [&](int start, std::vector<int> &order, std::vector<int> &lvl_ptr, int stamp_base) {
        // levels written as stamp_base + level into `level` (re-used across sweeps via negative marks)
        order.clear(); lvl_ptr.clear();
        int head = 0, tail = 0;
        queue[(size_t)tail++] = start;
        level[(size_t)start] = stamp_base;
        lvl_ptr.push_back(0);
        int cur = stamp_base;
        while (head < tail) {
            int v = queue[(size_t)head++];
            if (level[(size_t)v] != cur) { cur = level[(size_t)v]; lvl_ptr.push_back((int)order.size()); }
            order.push_back(v);
            for (int64_t p = Ap[(size_t)v]; p < Ap[(size_t)v + 1]; p++) {
                int u = Ai[(size_t)p];
                if (level[(size_t)u] < stamp_base) { level[(size_t)u] = level[(size_t)v] + 1; queue[(size_t)tail++] = u; }
            }
        }
        lvl_ptr.push_back((int)order.size());
    };
    auto order_interior = [&](const std::vector<int> &nodes) {
        // minimum degree on the induced subgraph
        const int k = (int)nodes.size();
        if (k == 0) return;
        for (int i = 0; i < k; i++) local[(size_t)nodes[(size_t)i]] = i;
        std::vector<int64_t> sp((size_t)k + 1, 0);
        std::vector<int> si;
        for (int i = 0; i < k; i++) {
            int v = nodes[(size_t)i];
            for (int64_t p = Ap[(size_t)v]; p < Ap[(size_t)v + 1]; p++) {
                int u = local[(size_t)Ai[(size_t)p]];
                if (u >= 0) si.push_back(u);
            }
            sp[(size_t)i + 1] = (int64_t)si.size();
        }
        std::vector<int> lp;
        amd_order(k, sp, si, lp);
        for (int i = 0; i < k; i++) Pout.push_back(nodes[(size_t)lp[(size_t)i]]);
        for (int i = 0; i < k; i++) local[(size_t)nodes[(size_t)i]] = -1;
    };
    int stamp = 0;
    std::vector<int> order, lvl_ptr;
    for (int root = 0; root < n; root++) {
        if (visited[(size_t)root]) continue;
        // pseudo-peripheral start: a few BFS sweeps, restarting from a min-degree node of the last level
        int start = root;
        int nlev = 0;
        for (int sweep = 0; sweep < 4; sweep++) {
            stamp += n + 2;
            bfs(start, order, lvl_ptr, stamp);
            int nl = (int)lvl_ptr.size() - 1;
            if (sweep > 0 && nl <= nlev) break;
            nlev = nl;
            int best = -1;
            int64_t bestdeg = INT64_MAX;
            for (int q = lvl_ptr[(size_t)nl - 1]; q < lvl_ptr[(size_t)nl]; q++) {
                int v = order[(size_t)q];
                int64_t d = Ap[(size_t)v + 1] - Ap[(size_t)v];
                if (d < bestdeg) { bestdeg = d; best = v; }
            }
            if (best < 0 || best == start) break;
            if (sweep < 3) start = best;
        }
        // final level structure from `start`
        stamp += n + 2;
        bfs(start, order, lvl_ptr, stamp);
        const int nc = (int)order.size();
        const int nl = (int)lvl_ptr.size() - 1;
        for (int v : order) visited[(size_t)v] = 1;
        int p = std::max(1, nc / part_size);
        std::vector<int> seps;     // separator levels (ascending)
        if (p >= 2 && nl >= 5) {
            int last = 0;
            for (int k = 1; k < p; k++) {
                const int64_t target = (int64_t)nc * k / p;
                int l = (int)(std::upper_bound(lvl_ptr.begin(), lvl_ptr.end(), (int)target) - lvl_ptr.begin()) - 1;
                l = std::max(l, last + 2);
                if (l >= nl - 1) break;
                // the thinnest of the neighbouring levels
                int bestl = l;
                for (int c = l - 1; c <= l + 1; c++) {
                    if (c < last + 2 || c >= nl - 1) continue;
                    if (lvl_ptr[(size_t)c + 1] - lvl_ptr[(size_t)c] < lvl_ptr[(size_t)bestl + 1] - lvl_ptr[(size_t)bestl]) bestl = c;
                }
                seps.push_back(bestl);
                last = bestl;
            }
        }
        // interiors
        int lo = 0;
        std::vector<int> nodes;
        for (size_t k = 0; k <= seps.size(); k++) {
            const int hi = (k < seps.size()) ? seps[k] : nl;     // levels [lo, hi)
            nodes.assign(order.begin() + lvl_ptr[(size_t)lo], order.begin() + lvl_ptr[(size_t)hi]);
            order_interior(nodes);
            lo = hi + 1;
        }
        // separators of this component, ranked like cyclic reduction
        for (size_t k = 0; k < seps.size(); k++) {
            const int l = seps[k];
            std::vector<int> sn(order.begin() + lvl_ptr[(size_t)l], order.begin() + lvl_ptr[(size_t)l + 1]);
            int idx = (int)k + 1, tz = 0;
            while ((idx & 1) == 0) { idx >>= 1; tz++; }
            sep_rank.push_back({tz, (int)separators_all.size()});
            separators_all.push_back(std::move(sn));
        }
    }
    std::stable_sort(sep_rank.begin(), sep_rank.end(),
                     [](const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.first < b.first; });
    for (auto &r : sep_rank)
        for (int v : separators_all[(size_t)r.second]) Pout.push_back(v);
    if ((int)Pout.size() != n) throw std::runtime_error("dissection_order: internal error");
}

// ------------------------------------------------------------------------------------------------
static bool relax_ok(int W, int64_t stored, int64_t zeros) {
    double z = stored > 0 ? (double)zeros / (double)stored : 0.0;
    if (W <= 4) return true;
    if (W <= 16) return z < 0.5;
    if (W <= 48) return z < 0.12;
    return z < 0.05;
}

// symmetric pattern of K without the diagonal (variables <-> constraints), sorted, de-duplicated
void build_kkt_graph(int nvar, int ncon, int64_t nnzj, const int64_t *jrow, const int64_t *jcol,
                     std::vector<int64_t> &Gp, std::vector<int> &Gi) {
    const int N = nvar + ncon;
    Gp.assign((size_t)N + 1, 0);
    for (int64_t e = 0; e < nnzj; e++) { Gp[(size_t)jcol[e] + 1]++; Gp[(size_t)(nvar + jrow[e]) + 1]++; }
    for (int i = 0; i < N; i++) Gp[i + 1] += Gp[i];
    Gi.assign((size_t)Gp[N], 0);
    std::vector<int64_t> pos(Gp.begin(), Gp.end() - 1);
    for (int64_t e = 0; e < nnzj; e++) {
        int v = (int)jcol[e], c = nvar + (int)jrow[e];
        Gi[(size_t)pos[v]++] = c;
        Gi[(size_t)pos[c]++] = v;
    }
    std::vector<int64_t> Np((size_t)N + 1, 0);
    int64_t out = 0;
    for (int i = 0; i < N; i++) {
        int64_t a = Gp[i], b = Gp[i + 1];
        std::sort(Gi.begin() + a, Gi.begin() + b);
        int64_t start = out;
        for (int64_t p = a; p < b; p++)
            if (out == start || Gi[(size_t)(out - 1)] != Gi[(size_t)p]) Gi[(size_t)out++] = Gi[(size_t)p];
        Np[i + 1] = out;
    }
    Gp = Np;
    Gi.resize((size_t)out);
}

constexpr int kMaxSuperWidth = 64;

void analyze(int nvar, int ncon, int64_t nnzj, const int64_t *jrow, const int64_t *jcol,
             const int64_t *Puser, Symbolic &S) {
    const int N = nvar + ncon;
    S = Symbolic();
    S.N = N;
    std::vector<int64_t> Gp;
    std::vector<int> Gi;
    build_kkt_graph(nvar, ncon, nnzj, jrow, jcol, Gp, Gi);
    // ---- ordering ---------------------------------------------------------------------------
    S.P.resize((size_t)N);
    if (Puser) {
        std::vector<char> seen((size_t)N, 0);
        for (int k = 0; k < N; k++) {
            int64_t v = Puser[k];
            if (v < 0 || v >= N || seen[(size_t)v]) throw std::invalid_argument("P is not a permutation");
            seen[(size_t)v] = 1;
            S.P[(size_t)k] = (int)v;
        }
    } else {
        amd_order(N, Gp, Gi, S.P);
    }
    S.pinv.resize((size_t)N);
    for (int k = 0; k < N; k++) S.pinv[(size_t)S.P[(size_t)k]] = k;
    const std::vector<int> &pinv = S.pinv;
    // ---- permuted strict-lower pattern by columns (Bp/Bi) and by rows (Up/Ui) -------------------
    std::vector<int64_t> Bp((size_t)N + 1, 0), Up((size_t)N + 1, 0);
    for (int i = 0; i < N; i++) {
        int pi = pinv[(size_t)i];
        for (int64_t p = Gp[i]; p < Gp[i + 1]; p++) {
            int pj = pinv[(size_t)Gi[(size_t)p]];
            if (pj > pi) { Bp[(size_t)pi + 1]++; Up[(size_t)pj + 1]++; }
        }
    }
    for (int i = 0; i < N; i++) { Bp[i + 1] += Bp[i]; Up[i + 1] += Up[i]; }
    std::vector<int> Bi((size_t)Bp[N]), Ui((size_t)Up[N]);
    {
        std::vector<int64_t> bpos(Bp.begin(), Bp.end() - 1), upos(Up.begin(), Up.end() - 1);
        for (int k = 0; k < N; k++) {          // ascending permuted column => Ui rows sorted by column
            int i = S.P[(size_t)k];
            for (int64_t p = Gp[i]; p < Gp[i + 1]; p++) {
                int pj = pinv[(size_t)Gi[(size_t)p]];
                if (pj > k) { Bi[(size_t)bpos[k]++] = pj; Ui[(size_t)upos[pj]++] = k; }
            }
        }
        for (int k = 0; k < N; k++) std::sort(Bi.begin() + Bp[k], Bi.begin() + Bp[k + 1]);
    }
    // ---- elimination tree (Liu) ----------------------------------------------------------------
    S.parent.assign((size_t)N, -1);
    {
        std::vector<int> anc((size_t)N, -1);
        for (int i = 0; i < N; i++) {
            for (int64_t p = Up[i]; p < Up[i + 1]; p++) {
                int j = Ui[(size_t)p];
                while (j != -1 && j < i) {
                    int nx = anc[(size_t)j];
                    anc[(size_t)j] = i;
                    if (nx == -1) { S.parent[(size_t)j] = i; }
                    j = nx;
                }
            }
        }
    }
    // ---- column structures of L by child merging ------------------------------------------------
    S.Lp.assign((size_t)N + 1, 0);
    {
        std::vector<int> chead((size_t)N, -1), cnext((size_t)N, -1);
        for (int j = N - 1; j >= 0; j--) {
            int pa = S.parent[(size_t)j];
            if (pa >= 0) { cnext[(size_t)j] = chead[(size_t)pa]; chead[(size_t)pa] = j; }
        }
        std::vector<int> mark((size_t)N, -1);
        std::vector<int> tmp;
        S.Li.clear();
        S.Li.reserve((size_t)(Bp[N] * 3));
        for (int j = 0; j < N; j++) {
            tmp.clear();
            mark[(size_t)j] = j;
            for (int64_t p = Bp[j]; p < Bp[j + 1]; p++) {
                int r = Bi[(size_t)p];
                if (mark[(size_t)r] != j) { mark[(size_t)r] = j; tmp.push_back(r); }
            }
            for (int c = chead[(size_t)j]; c != -1; c = cnext[(size_t)c]) {
                for (int64_t p = S.Lp[(size_t)c]; p < S.Lp[(size_t)c + 1]; p++) {
                    int r = S.Li[(size_t)p];
                    if (mark[(size_t)r] != j) { mark[(size_t)r] = j; tmp.push_back(r); }
                }
            }
            std::sort(tmp.begin(), tmp.end());
            S.Li.insert(S.Li.end(), tmp.begin(), tmp.end());
            S.Lp[(size_t)j + 1] = (int64_t)S.Li.size();
        }
    }
    auto cnt = [&](int j) { return (int)(S.Lp[(size_t)j + 1] - S.Lp[(size_t)j]); };
    // ---- supernodes: exact nesting, then relaxed amalgamation, width cap ------------------------
    std::vector<int> first;   // first columns
    {
        // pass 1: fundamental-style chains
        std::vector<int> f1;
        f1.push_back(0);
        for (int j = 0; j + 1 < N; j++) {
            bool join = (S.parent[(size_t)j] == j + 1) && (cnt(j + 1) == cnt(j) - 1);
            if (!join) f1.push_back(j + 1);
        }
        f1.push_back(N);
        // pass 2: relaxed merges of consecutive chains + width cap
        first.push_back(0);
        int cur_f = 0;                    // current merged supernode [cur_f, cur_l]
        int64_t cur_true = 0;
        for (size_t s = 0; s + 1 < f1.size(); s++) {
            int f = f1[s], l = f1[s + 1] - 1;
            int64_t tr = 0;
            for (int j = f; j <= l; j++) tr += cnt(j) + 1;
            if (s == 0) { cur_f = f; cur_true = tr; continue; }
            int prev_l = f - 1;
            bool can = (S.parent[(size_t)prev_l] == f);
            if (can) {
                int W = l - cur_f + 1;
                int nr = cnt(l);
                int64_t stored = (int64_t)W * (W + 1) / 2 + (int64_t)W * nr;
                int64_t zeros = stored - (cur_true + tr);
                can = (W <= kMaxSuperWidth) && relax_ok(W, stored, zeros);
            }
            if (can) { cur_true += tr; }
            else { first.push_back(f); cur_f = f; cur_true = tr; }
        }
        if (N > 0) first.push_back(N);
        // width cap: split wide supernodes into panels of <= kMaxSuperWidth columns
        std::vector<int> f2;
        for (size_t s = 0; s + 1 < first.size(); s++) {
            int f = first[s], l = first[s + 1];
            for (int c = f; c < l; c += kMaxSuperWidth) f2.push_back(c);
        }
        f2.push_back(N);
        first.swap(f2);
        if (N == 0) first.assign(1, 0);
    }
    S.nsuper = (int)first.size() - 1;
    S.sfirst = first;
    S.sn_of.assign((size_t)N, 0);
    for (int s = 0; s < S.nsuper; s++)
        for (int j = first[(size_t)s]; j < first[(size_t)s + 1]; j++) S.sn_of[(size_t)j] = s;
    // row structures: R_s = struct(last column of s) (merged / split panels included: the last
    // column's structure is exactly the set of rows below the block)
    S.rptr.assign((size_t)S.nsuper + 1, 0);
    S.poff.assign((size_t)S.nsuper + 1, 0);
    for (int s = 0; s < S.nsuper; s++) {
        int l = first[(size_t)s + 1] - 1;
        int w = first[(size_t)s + 1] - first[(size_t)s];
        int nr = cnt(l);
        S.rptr[(size_t)s + 1] = S.rptr[(size_t)s] + nr;
        S.poff[(size_t)s + 1] = S.poff[(size_t)s] + (int64_t)w * (w + nr);
    }
    S.rows.resize((size_t)S.rptr[(size_t)S.nsuper]);
    for (int s = 0; s < S.nsuper; s++) {
        int l = first[(size_t)s + 1] - 1;
        std::copy(S.Li.begin() + S.Lp[(size_t)l], S.Li.begin() + S.Lp[(size_t)l + 1],
                  S.rows.begin() + S.rptr[(size_t)s]);
    }
    S.panel_size = S.poff[(size_t)S.nsuper];
    // ---- update pairs (source d -> target t), relative indices, levels --------------------------
    struct Pair { int t, d, a, b; };
    std::vector<Pair> pairs;
    S.level.assign((size_t)S.nsuper, 0);
    S.tptr.assign((size_t)S.nsuper + 1, 0);
    for (int d = 0; d < S.nsuper; d++) {
        int64_t r0 = S.rptr[(size_t)d], r1 = S.rptr[(size_t)d + 1];
        int nr = (int)(r1 - r0);
        int a = 0;
        while (a < nr) {
            int t = S.sn_of[(size_t)S.rows[(size_t)(r0 + a)]];
            int b = a + 1;
            while (b < nr && S.sn_of[(size_t)S.rows[(size_t)(r0 + b)]] == t) b++;
            pairs.push_back({t, d, a, b});
            S.level[(size_t)t] = std::max(S.level[(size_t)t], S.level[(size_t)d] + 1);
            a = b;
        }
    }
    // by-source list (pairs are generated in ascending d, ascending t)
    S.ttgt.resize(pairs.size());
    for (size_t k = 0; k < pairs.size(); k++) { S.tptr[(size_t)pairs[k].d + 1]++; S.ttgt[k] = pairs[k].t; }
    for (int s = 0; s < S.nsuper; s++) S.tptr[(size_t)s + 1] += S.tptr[(size_t)s];
    // by-target list: non-leaf sources first, then leaf sources; ascending d inside each group
    S.uptr.assign((size_t)S.nsuper + 1, 0);
    for (auto &p : pairs) S.uptr[(size_t)p.t + 1]++;
    for (int s = 0; s < S.nsuper; s++) S.uptr[(size_t)s + 1] += S.uptr[(size_t)s];
    S.umid.assign((size_t)S.nsuper, 0);
    S.usrc.resize(pairs.size()); S.ua.resize(pairs.size()); S.ub.resize(pairs.size());
    S.urel.resize(pairs.size());
    {
        std::vector<int64_t> pos(S.uptr.begin(), S.uptr.end() - 1);
        for (int pass = 0; pass < 2; pass++) {
            for (size_t k = 0; k < pairs.size(); k++) {
                const bool leaf = S.level[(size_t)pairs[k].d] == 0;
                if ((pass == 0) == leaf) continue;
                int64_t q = pos[(size_t)pairs[k].t]++;
                S.usrc[(size_t)q] = pairs[k].d; S.ua[(size_t)q] = pairs[k].a; S.ub[(size_t)q] = pairs[k].b;
            }
            if (pass == 0) for (int t = 0; t < S.nsuper; t++) S.umid[(size_t)t] = pos[(size_t)t];
        }
        int64_t relsz = 0;
        for (size_t q = 0; q < pairs.size(); q++) {
            int d = S.usrc[q];
            int nr = (int)(S.rptr[(size_t)d + 1] - S.rptr[(size_t)d]);
            S.urel[q] = relsz;
            relsz += nr - S.ua[q];
        }
        S.rel.resize((size_t)relsz);
        for (int t = 0; t < S.nsuper; t++) {
            int ft = first[(size_t)t], lt = first[(size_t)t + 1] - 1, wt = lt - ft + 1;
            const int *Rt = S.rows.data() + S.rptr[(size_t)t];
            int nrt = (int)(S.rptr[(size_t)t + 1] - S.rptr[(size_t)t]);
            for (int64_t q = S.uptr[(size_t)t]; q < S.uptr[(size_t)t + 1]; q++) {
                int d = S.usrc[(size_t)q];
                const int *Rd = S.rows.data() + S.rptr[(size_t)d];
                int nrd = (int)(S.rptr[(size_t)d + 1] - S.rptr[(size_t)d]);
                int *out = S.rel.data() + S.urel[(size_t)q];
                int cursor = 0;
                for (int i = S.ua[(size_t)q]; i < nrd; i++) {
                    int r = Rd[i];
                    if (r <= lt) { out[i - S.ua[(size_t)q]] = r - ft; }
                    else {
                        while (cursor < nrt && Rt[cursor] < r) cursor++;
                        if (cursor >= nrt || Rt[cursor] != r)
                            throw std::runtime_error("symbolic: update row missing from target structure");
                        out[i - S.ua[(size_t)q]] = wt + cursor;
                    }
                }
            }
        }
        // leaf contributions grouped by target column (counting sort by column; ascending d kept
        // because targets are visited in ascending t and their leaf pairs in ascending d... the
        // order inside a column is made explicit by sorting on d below)
        S.lcptr.assign((size_t)N + 1, 0);
        for (int t = 0; t < S.nsuper; t++)
            for (int64_t q = S.umid[(size_t)t]; q < S.uptr[(size_t)t + 1]; q++) {
                const int d = S.usrc[(size_t)q];
                const int *Rd = S.rows.data() + S.rptr[(size_t)d];
                for (int jp = S.ua[(size_t)q]; jp < S.ub[(size_t)q]; jp++) S.lcptr[(size_t)Rd[jp] + 1]++;
            }
        for (int j = 0; j < N; j++) S.lcptr[(size_t)j + 1] += S.lcptr[(size_t)j];
        const size_t nlc = (size_t)S.lcptr[(size_t)N];
        S.lc_src.resize(nlc); S.lc_rel.resize(nlc); S.lc_cnt.resize(nlc);
        S.lc_ldd.resize(nlc); S.lc_wd.resize(nlc); S.lc_fd.resize(nlc);
        std::vector<int64_t> cpos(S.lcptr.begin(), S.lcptr.end() - 1);
        for (int t = 0; t < S.nsuper; t++)
            for (int64_t q = S.umid[(size_t)t]; q < S.uptr[(size_t)t + 1]; q++) {
                const int d = S.usrc[(size_t)q];
                const int *Rd = S.rows.data() + S.rptr[(size_t)d];
                const int fd = first[(size_t)d], wd = first[(size_t)d + 1] - fd;
                const int nrd = (int)(S.rptr[(size_t)d + 1] - S.rptr[(size_t)d]);
                for (int jp = S.ua[(size_t)q]; jp < S.ub[(size_t)q]; jp++) {
                    const size_t e = (size_t)cpos[(size_t)Rd[jp]]++;
                    S.lc_src[e] = S.poff[(size_t)d] + wd + jp;
                    S.lc_rel[e] = S.urel[(size_t)q] + (jp - S.ua[(size_t)q]);
                    S.lc_cnt[e] = nrd - jp;
                    S.lc_ldd[e] = wd + nrd;
                    S.lc_wd[e] = wd;
                    S.lc_fd[e] = fd;
                }
            }
    }
    S.nlevels = 0; S.nleaf = 0;
    for (int t = 0; t < S.nsuper; t++) {
        S.nlevels = std::max(S.nlevels, S.level[(size_t)t] + 1);
        if (S.level[(size_t)t] == 0) S.nleaf++;
    }
    // task order: by (level, index)
    S.order.resize((size_t)S.nsuper);
    std::iota(S.order.begin(), S.order.end(), 0);
    std::stable_sort(S.order.begin(), S.order.end(),
                     [&](int x, int y) { return S.level[(size_t)x] < S.level[(size_t)y]; });
    // ---- assembly map ------------------------------------------------------------------------
    {
        const int64_t nent = (int64_t)nvar + nnzj + ncon;
        std::vector<std::pair<int64_t, int>> ent((size_t)nent);
        auto slot = [&](int pi, int pj) -> int64_t {
            int col = std::min(pi, pj), row = std::max(pi, pj);
            int s = S.sn_of[(size_t)col];
            int fs = first[(size_t)s], ls = first[(size_t)s + 1] - 1, w = ls - fs + 1;
            int nr = (int)(S.rptr[(size_t)s + 1] - S.rptr[(size_t)s]);
            int ld = w + nr;
            int lr;
            if (row <= ls) lr = row - fs;
            else {
                const int *R = S.rows.data() + S.rptr[(size_t)s];
                const int *it = std::lower_bound(R, R + nr, row);
                if (it == R + nr || *it != row) throw std::runtime_error("symbolic: entry outside the fill pattern");
                lr = w + (int)(it - R);
            }
            return S.poff[(size_t)s] + (int64_t)(col - fs) * ld + lr;
        };
        for (int k = 0; k < nvar; k++) ent[(size_t)k] = {slot(pinv[(size_t)k], pinv[(size_t)k]), k};
        for (int64_t e = 0; e < nnzj; e++)
            ent[(size_t)(nvar + e)] = {slot(pinv[(size_t)jcol[e]], pinv[(size_t)(nvar + jrow[e])]), (int)(nvar + e)};
        for (int i = 0; i < ncon; i++)
            ent[(size_t)(nvar + nnzj + i)] = {slot(pinv[(size_t)(nvar + i)], pinv[(size_t)(nvar + i)]), (int)(nvar + nnzj + i)};
        std::sort(ent.begin(), ent.end());
        S.aslot.clear(); S.aptr.clear(); S.asrc.resize((size_t)nent);
        for (int64_t k = 0; k < nent; k++) {
            if (k == 0 || ent[(size_t)k].first != ent[(size_t)k - 1].first) {
                S.aslot.push_back(ent[(size_t)k].first);
                S.aptr.push_back(k);
            }
            S.asrc[(size_t)k] = ent[(size_t)k].second;
        }
        S.aptr.push_back(nent);
    }
    // flops of the numeric factorisation (2 * sum_j nnz(L_j)^2)
    S.flops = 0;
    for (int j = 0; j < N; j++) { double c = (double)cnt(j); S.flops += 2.0 * c * c; }
}

}  // namespace fpsb
