// fpsb_symbolic.h — host-side symbolic analysis of K = [I A'; A -dI] (ldl_analyze equivalent)
#pragma once
#include <stdint.h>
#include <vector>

namespace fpsb {

struct Symbolic {
    int N = 0;
    std::vector<int> P, pinv;         // P[k] = original index eliminated k-th
    std::vector<int> parent;          // elimination tree (-1 = root)
    std::vector<int64_t> Lp;          // N+1, strict lower column pointers
    std::vector<int> Li;              // row indices (ascending per column)
    // supernodal plan
    int nsuper = 0;
    std::vector<int> sfirst;          // nsuper+1 : first column of each supernode
    std::vector<int> sn_of;           // N : supernode of each column
    std::vector<int64_t> rptr;        // nsuper+1 : offsets into rows[]
    std::vector<int> rows;            // below-block row structure of each supernode (ascending)
    std::vector<int64_t> poff;        // nsuper+1 : panel offsets (doubles); ld = w + nr
    std::vector<int> level;           // nsuper : height from the leaves of the supernodal tree
    std::vector<int> order;           // nsuper : supernodes sorted by (level, index) — task order
    // update pairs, grouped by target (ascending source): source d contributes rows[a..b) of d
    std::vector<int64_t> uptr;        // nsuper+1
    std::vector<int64_t> umid;        // nsuper : pairs [uptr[t], umid[t]) have non-leaf sources,
                                      //          [umid[t], uptr[t+1]) have leaf (level-0) sources
    std::vector<int> usrc, ua, ub;    // per pair
    std::vector<int64_t> urel;        // per pair: offset into rel[]
    std::vector<int> rel;             // local row index in the target panel of d's rows a..nr_d
    // pairs grouped by source (for the backward solve dependency waits): targets of each source
    std::vector<int64_t> tptr;        // nsuper+1
    std::vector<int> ttgt;            // distinct target supernodes of each source
    // leaf contributions by target column: for permuted column j, entries [lcptr[j], lcptr[j+1])
    // describe the leaf supernodes d whose row structure contains j (ascending d)
    std::vector<int64_t> lcptr;       // N+1
    std::vector<int64_t> lc_src;      // panel offset of L_d(jp, 0)
    std::vector<int64_t> lc_rel;      // offset into rel[] of the local row index of row jp
    std::vector<int> lc_cnt;          // rows jp..nr_d-1 of d
    std::vector<int> lc_ldd, lc_wd, lc_fd;
    // assembly map: panel slot <- sources (index into [1..1 | jvals | -delta..])
    std::vector<int64_t> aslot;       // distinct target slots (panel offsets)
    std::vector<int64_t> aptr;        // naslot+1
    std::vector<int> asrc;            // source ids
    int64_t panel_size = 0;
    double flops = 0;
    int nlevels = 0;                  // height of the supernodal dependency DAG
    int nleaf = 0;                    // supernodes with no incoming update (level 0)
};

// approximate minimum degree ordering of a symmetric pattern (Ap/Ai: full pattern, no diagonal)
void amd_order(int n, const std::vector<int64_t> &Ap, const std::vector<int> &Ai, std::vector<int> &P);

// BFS level-set dissection + cyclic-reduction separator order (nparts_target <= 0: automatic)
void dissection_order(int n, const std::vector<int64_t> &Ap, const std::vector<int> &Ai, int nparts_target,
                      std::vector<int> &P);
void build_kkt_graph(int nvar, int ncon, int64_t nnzj, const int64_t *jrow, const int64_t *jcol,
                     std::vector<int64_t> &Gp, std::vector<int> &Gi);

// full analysis; Puser may be null (-> amd_order). jrow/jcol are 0-based.
void analyze(int nvar, int ncon, int64_t nnzj, const int64_t *jrow, const int64_t *jcol,
             const int64_t *Puser, Symbolic &S);

}  // namespace fpsb
