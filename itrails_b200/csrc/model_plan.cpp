// model_plan.cpp — see model_plan.h.
#include "model_plan.h"

#include <algorithm>
#include <functional>
#include <stdexcept>
#include <tuple>

namespace itr {

// ---------------------------------------------------------------------------------
// state spaces                                         trans_mat.py:26-194, 269-286
// ---------------------------------------------------------------------------------
Labels StateSpace::canon(const int *labels, int m) {
    int map[16];
    for (int &v : map) v = 0;
    int next = 0;
    Labels out{};
    for (int k = 0; k < m; ++k) {
        if (!map[labels[k]]) map[labels[k]] = ++next;
        out[k] = (uint8_t)map[labels[k]];
    }
    return out;
}

void StateSpace::build(int n_species) {
    n = n_species;
    const int m = 2 * n;
    states.clear();
    trans.clear();
    index.clear();
    // set partitions of the 2n lineage ends as restricted-growth strings
    std::function<void(Labels &, int, int)> rec = [&](Labels &cur, int pos, int mx) {
        if (pos == m) {
            states.push_back(cur);
            return;
        }
        for (int lab = 1; lab <= mx + 1; ++lab) {
            cur[pos] = (uint8_t)lab;
            rec(cur, pos + 1, std::max(mx, lab));
        }
        cur[pos] = 0;
    };
    Labels cur{};
    rec(cur, 0, 0);
    size = (int)states.size();
    for (int i = 0; i < size; ++i) index[states[i]] = i;

    for (int i = 0; i < size; ++i) {
        const Labels &s = states[i];
        int mx = 0;
        for (int k = 0; k < m; ++k) mx = std::max<int>(mx, s[k]);
        // coalescence of two blocks (trans_mat.py:74-133)
        for (int x = 1; x <= mx; ++x)
            for (int y = x + 1; y <= mx; ++y) {
                int t[6];
                for (int k = 0; k < m; ++k) t[k] = (s[k] == y) ? x : s[k];
                trans.push_back({i, index.at(canon(t, m)), 1});
            }
        // recombination of a block that holds both a left and a right end
        // (trans_mat.py:134-194)
        for (int x = 1; x <= mx; ++x) {
            bool left = false, right = false;
            for (int k = 0; k < m; ++k)
                if (s[k] == x) (k < n ? left : right) = true;
            if (!(left && right)) continue;
            int t[6];
            for (int k = 0; k < m; ++k) t[k] = (s[k] == x && k >= n) ? mx + 1 : s[k];
            trans.push_back({i, index.at(canon(t, m)), 2});
        }
    }
    omega_l.assign(size, 0);
    omega_r.assign(size, 0);
    for (int i = 0; i < size; ++i) {
        const Labels &s = states[i];
        for (int half = 0; half < 2; ++half) {
            int w = 0;
            for (int k = 0; k < n; ++k) {
                int cnt = 0;
                for (int q = 0; q < n; ++q) cnt += s[half * n + q] == s[half * n + k];
                if (cnt > 1) w |= 1 << k;
            }
            (half ? omega_r : omega_l)[i] = w;
        }
    }
}

// combine_states.py:5-80: index of the merged state for every pair of states of two
// independent chains (ends ordered left_1, left_2, right_1, right_2).
static std::vector<int> combine_map(const StateSpace &a, const StateSpace &b, const StateSpace &ab) {
    std::vector<int> out((size_t)a.size * b.size);
    const int n1 = a.n, n2 = b.n;
    for (int i1 = 0; i1 < a.size; ++i1) {
        int off = 0;
        for (int k = 0; k < 2 * n1; ++k) off = std::max<int>(off, a.states[i1][k]);
        for (int i2 = 0; i2 < b.size; ++i2) {
            int merged[6], p = 0;
            for (int k = 0; k < n1; ++k) merged[p++] = a.states[i1][k];
            for (int k = 0; k < n2; ++k) merged[p++] = b.states[i2][k] + off;
            for (int k = 0; k < n1; ++k) merged[p++] = a.states[i1][n1 + k];
            for (int k = 0; k < n2; ++k) merged[p++] = b.states[i2][n2 + k] + off;
            out[(size_t)i1 * b.size + i2] = ab.index.at(StateSpace::canon(merged, p));
        }
    }
    return out;
}

static GenCSR restricted_generator(const StateSpace &ss, const std::vector<int> &subset) {
    GenCSR g;
    g.n = (int)subset.size();
    std::vector<int> local(ss.size, -1);
    for (int k = 0; k < g.n; ++k) local[subset[k]] = k;
    std::vector<std::vector<std::pair<int, int>>> rows(g.n);
    g.ncoal.assign(g.n, 0);
    g.nrec.assign(g.n, 0);
    for (const auto &t : ss.trans) {
        const int f = local[t[0]];
        if (f < 0) continue;
        (t[2] == 2 ? g.nrec : g.ncoal)[f] += 1;          // the diagonal keeps every exit
        if (local[t[1]] >= 0) rows[f].push_back({local[t[1]], t[2]});
    }
    g.row_ptr.assign(g.n + 1, 0);
    for (int r = 0; r < g.n; ++r) {
        std::sort(rows[r].begin(), rows[r].end());
        g.row_ptr[r + 1] = g.row_ptr[r] + (int)rows[r].size();
        for (auto &e : rows[r]) {
            g.col.push_back(e.first);
            g.kind.push_back((uint8_t)e.second);
        }
    }
    g.transient.assign(g.n, 0);
    for (int k = 0; k < g.n; ++k)
        g.transient[k] = (ss.omega_l[subset[k]] == 0 || ss.omega_r[subset[k]] == 0) ? 1 : 0;
    return g;
}

// ---------------------------------------------------------------------------------
// per-locus genealogy histories (the reference's path keys)
//   (-1,-1,-1) nothing yet | (0,i,-1) A,B coalesced in AB interval i | (k,s,-1) first
//   coalescence in ABC interval s with topology k | (k,s,u) second coalescence in u
// ---------------------------------------------------------------------------------
using Hist = std::array<int, 3>;
static const Hist NONE_H = {-1, -1, -1};
static int class_of_topo(int k) { return k == 2 ? 5 : k == 3 ? 6 : 3; }   // topologies 0 and 1 -> {A,B}
static int hist_class(const Hist &h) {                                     // helper_omegas.py:25-87
    if (h[0] == -1) return 0;
    return h[2] != -1 ? 7 : class_of_topo(h[0]);
}
static const int FIRSTS[3] = {3, 5, 6};
static int topo_of_class(int c) { return c == 3 ? 1 : c == 5 ? 2 : 3; }
static int first_id(int c) { return c == 3 ? 0 : c == 5 ? 1 : 2; }

struct Succ {
    Hist h;
    int x;   // constraining first-coalescence class, 0 = unconstrained
};

// run_markov_chain_ABC.py:368-392 + vanloan.py:366-371
static std::vector<Succ> succ_abc(const Hist &h, int s) {
    std::vector<Succ> out;
    if (h[0] == -1) {
        out.push_back({h, 0});
        for (int x : FIRSTS) {
            out.push_back({{topo_of_class(x), s, -1}, x});
            out.push_back({{topo_of_class(x), s, s}, x});
        }
    } else if (h[2] == -1) {
        const int x = class_of_topo(h[0]);
        out.push_back({h, x});
        out.push_back({{h[0], h[1], s}, x});
    } else {
        out.push_back({h, 0});
    }
    return out;
}

// run_markov_chain_ABC.py:519-795 (last, unbounded interval)
static std::vector<Succ> finals(const Hist &h, int last) {
    std::vector<Succ> out;
    if (h[0] == -1) {
        for (int x : FIRSTS) out.push_back({{topo_of_class(x), last, last}, x});
    } else if (h[2] == -1) {
        out.push_back({{h[0], h[1], last}, class_of_topo(h[0])});
    } else {
        out.push_back({h, 0});
    }
    return out;
}

using Key = std::pair<Hist, Hist>;
struct KeyInfo {
    int off;   // offset of the key's vector (compact over its class) in the stage buffer
    int cl, cr;
};

void ModelPlan::build(int n_ab, int n_abc) {
    if (n_ab < 1 || n_abc < 1) throw std::runtime_error("n_int_AB and n_int_ABC must be >= 1");
    n_int_AB = n_ab;
    n_int_ABC = n_abc;
    ss1.build(1);
    ss2.build(2);
    ss3.build(3);
    if (ss1.size != 2 || ss2.size != 15 || ss3.size != 203) throw std::runtime_error("state-space sizes");
    if (ss1.states[0][0] != 1 || ss1.states[0][1] != 1) throw std::runtime_error("state order of the one-sequence chain");

    // ---- generators --------------------------------------------------------------
    gens.clear();
    auto all_states = [](const StateSpace &ss) {
        std::vector<int> v(ss.size);
        for (int i = 0; i < ss.size; ++i) v[i] = i;
        return v;
    };
    gens.push_back(restricted_generator(ss1, all_states(ss1)));
    gens.push_back(restricted_generator(ss2, all_states(ss2)));
    std::vector<std::vector<int>> S(9);
    for (int xi = 0; xi < 3; ++xi)
        for (int yi = 0; yi < 3; ++yi) {
            const int x = FIRSTS[xi], y = FIRSTS[yi];
            std::vector<int> &set = S[xi * 3 + yi];
            for (int i = 0; i < ss3.size; ++i) {
                const int l = ss3.omega_l[i], r = ss3.omega_r[i];
                if ((l == 0 || l == x || l == 7) && (r == 0 || r == y || r == 7)) set.push_back(i);
            }
            if ((int)set.size() != 83) throw std::runtime_error("S_xy must have 83 states");
            gens.push_back(restricted_generator(ss3, set));
        }

    // ---- class lists ---------------------------------------------------------------
    std::map<std::pair<int, int>, std::vector<int>> cls2, cls3;
    for (int i = 0; i < ss2.size; ++i) cls2[{ss2.omega_l[i], ss2.omega_r[i]}].push_back(i);
    for (int i = 0; i < ss3.size; ++i) cls3[{ss3.omega_l[i], ss3.omega_r[i]}].push_back(i);
    idx_pool.clear();
    auto pool_add = [&](const std::vector<int> &v) {
        const int off = (int)idx_pool.size();
        idx_pool.insert(idx_pool.end(), v.begin(), v.end());
        return off;
    };
    // index lists are shared between ops: cache by (kind, a, b, c)
    std::map<std::tuple<int, int, int, int>, int> list_cache;
    auto cls2_list = [&](int l, int r) {
        auto key = std::make_tuple(0, l, r, 0);
        auto it = list_cache.find(key);
        if (it != list_cache.end()) return it->second;
        return list_cache[key] = pool_add(cls2.at({l, r}));
    };
    auto local_list = [&](int xy, int l, int r) {      // positions of class (l,r) inside S_xy
        auto key = std::make_tuple(1, xy, l, r);
        auto it = list_cache.find(key);
        if (it != list_cache.end()) return it->second;
        std::vector<int> pos;
        for (int st : cls3.at({l, r})) {
            auto p = std::lower_bound(S[xy].begin(), S[xy].end(), st);
            if (p == S[xy].end() || *p != st) throw std::runtime_error("class not inside S_xy");
            pos.push_back((int)(p - S[xy].begin()));
        }
        return list_cache[key] = pool_add(pos);
    };

    // ---- matrices ------------------------------------------------------------------
    // 0,1,2: one-sequence chain over t_A, t_B, t_C; 3..: AB intervals; then 9 per ABC interval
    mat_gen.clear();
    for (int k = 0; k < 3; ++k) mat_gen.push_back(0);
    for (int s = 0; s < n_ab; ++s) mat_gen.push_back(1);
    for (int s = 0; s + 1 < n_abc; ++s)
        for (int xy = 0; xy < 9; ++xy) mat_gen.push_back(2 + xy);
    n_mats = (int)mat_gen.size();
    mat_off.assign(n_mats, 0);
    mat_ld.assign(n_mats, 0);
    mat_pool = 0;
    for (int m = 0; m < n_mats; ++m) {
        const int np = gen_np(mat_gen[m]);
        mat_off[m] = (int32_t)mat_pool;
        mat_ld[m] = np;
        mat_pool += (int64_t)np * np;
    }
    auto mat_abc = [&](int s, int xy) { return 3 + n_ab + 9 * s + xy; };

    // ---- stages --------------------------------------------------------------------
    ops.clear();
    stages.clear();
    max_vec = 0;
    n_keys_max = 0;
    const std::vector<int> comb12 = combine_map(ss1, ss1, ss2);
    const std::vector<int> comb23 = combine_map(ss2, ss1, ss3);

    std::map<Key, KeyInfo> cur, nxt;
    int next_size = 0;
    auto begin_stage = [&]() {
        nxt.clear();
        next_size = 0;
        return (int)ops.size();
    };
    auto end_stage = [&](int op_begin, bool zero, bool final_stage = false) {
        stages.push_back({op_begin, (int)ops.size(), final_stage ? 0 : next_size, zero ? 1 : 0});
        max_vec = std::max(max_vec, next_size);
        n_keys_max = std::max<int64_t>(n_keys_max, (int64_t)nxt.size());
        cur.swap(nxt);
    };
    auto alloc_key = [&](const Key &k, int cl, int cr, int len) {
        if (nxt.count(k)) throw std::runtime_error("path key produced twice");
        KeyInfo ki{next_size, cl, cr};
        nxt[k] = ki;
        next_size += len;
        return ki.off;
    };

    // stage: product of the two one-sequence chains laid onto the two-sequence chain
    {
        const int ob = begin_stage();
        const std::vector<int> &c00 = cls2.at({0, 0});
        std::vector<int> scatter(4);
        for (int i1 = 0; i1 < 2; ++i1)
            for (int i2 = 0; i2 < 2; ++i2) {
                const int t = comb12[i1 * 2 + i2];
                auto p = std::find(c00.begin(), c00.end(), t);
                if (p == c00.end()) throw std::runtime_error("combined AB state outside class (0,0)");
                scatter[i1 * 2 + i2] = (int)(p - c00.begin());
            }
        const int dst = alloc_key({NONE_H, NONE_H}, 0, 0, (int)c00.size());
        ops.push_back({OP_INIT, 0, dst, 0, 0, pool_add(scatter), 2, 2});
        end_stage(ob, true);
    }
    // stages: AB intervals (run_markov_chain_AB.py:128-271)
    for (int s = 0; s < n_ab; ++s) {
        const int ob = begin_stage();
        for (const auto &kv : cur) {
            const Hist &hl = kv.first.first, &hr = kv.first.second;
            std::vector<Hist> ls{hl}, rs{hr};
            if (hl[0] == -1) ls.push_back({0, s, -1});
            if (hr[0] == -1) rs.push_back({0, s, -1});
            for (const Hist &hl2 : ls)
                for (const Hist &hr2 : rs) {
                    const int cl = hist_class(hl2), cr = hist_class(hr2);
                    const int len = (int)cls2.at({cl, cr}).size();
                    const int dst = alloc_key({hl2, hr2}, cl, cr, len);
                    ops.push_back({OP_MATVEC, kv.second.off, dst, 3 + s, cls2_list(kv.second.cl, kv.second.cr),
                                   cls2_list(cl, cr), (int)cls2.at({kv.second.cl, kv.second.cr}).size(), len});
                }
        }
        end_stage(ob, false);
    }
    // stage: merge with the C lineage (get_joint_prob_mat.py:155-161)
    {
        const int ob = begin_stage();
        for (const auto &kv : cur) {
            const int cl = kv.second.cl, cr = kv.second.cr;
            const std::vector<int> &src_states = cls2.at({cl, cr});
            const std::vector<int> &dst_states = cls3.at({cl, cr});
            std::vector<int> scatter;
            for (int st2 : src_states)
                for (int c = 0; c < 2; ++c) {
                    const int t = comb23[(size_t)st2 * 2 + c];
                    auto p = std::find(dst_states.begin(), dst_states.end(), t);
                    if (p == dst_states.end()) throw std::runtime_error("merged ABC state outside its class");
                    scatter.push_back((int)(p - dst_states.begin()));
                }
            auto ck = std::make_tuple(2, cl, cr, 0);
            int sc_off;
            auto it = list_cache.find(ck);
            if (it != list_cache.end()) sc_off = it->second;
            else sc_off = list_cache[ck] = pool_add(scatter);
            const int dst = alloc_key(kv.first, cl, cr, (int)dst_states.size());
            ops.push_back({OP_OUTER, kv.second.off, dst, 2, 0, sc_off, (int)src_states.size(), 2});
        }
        end_stage(ob, true);
    }
    // stages: bounded ABC intervals (run_markov_chain_ABC.py:347-518)
    for (int s = 0; s + 1 < n_abc; ++s) {
        const int ob = begin_stage();
        for (const auto &kv : cur) {
            const Hist &hl = kv.first.first, &hr = kv.first.second;
            const int c0l = kv.second.cl, c0r = kv.second.cr;
            for (const Succ &sl : succ_abc(hl, s))
                for (const Succ &sr : succ_abc(hr, s)) {
                    const int xy = first_id(sl.x ? sl.x : 3) * 3 + first_id(sr.x ? sr.x : 3);
                    const int c1l = hist_class(sl.h), c1r = hist_class(sr.h);
                    const int len = (int)cls3.at({c1l, c1r}).size();
                    const int dst = alloc_key({sl.h, sr.h}, c1l, c1r, len);
                    ops.push_back({OP_MATVEC, kv.second.off, dst, mat_abc(s, xy), local_list(xy, c0l, c0r),
                                   local_list(xy, c1l, c1r), (int)cls3.at({c0l, c0r}).size(), len});
                }
        }
        end_stage(ob, false);
    }

    // ---- hidden states (get_emission_prob_mat.py:803-1033 order, then sorted) --------
    hidden.clear();
    for (int i = 0; i < n_abc; ++i)
        for (int j = i + 1; j < n_abc; ++j)
            for (int k = 1; k <= 3; ++k) hidden.push_back({k, i, j});
    for (int i = 0; i < n_abc; ++i)
        for (int k = 1; k <= 3; ++k) hidden.push_back({k, i, i});
    for (int i = 0; i < n_ab; ++i)
        for (int j = 0; j < n_abc; ++j) hidden.push_back({0, i, j});
    std::sort(hidden.begin(), hidden.end(), [](const EmissionRecipe &a, const EmissionRecipe &b) {
        return std::tie(a.topo, a.i, a.j) < std::tie(b.topo, b.i, b.j);
    });
    K = (int)hidden.size();
    std::map<Hist, int> hidx;
    for (int k = 0; k < K; ++k) hidx[{hidden[k].topo, hidden[k].i, hidden[k].j}] = k;

    // stage: last interval -> joint matrix J (run_markov_chain_ABC.py:519-795)
    {
        const int ob = begin_stage();
        const int last = n_abc - 1;
        for (const auto &kv : cur) {
            const Hist &hl = kv.first.first, &hr = kv.first.second;
            const int len = (int)cls3.at({kv.second.cl, kv.second.cr}).size();
            if (hl[0] != -1 && hr[0] != -1) {
                const Hist hl2 = hl[2] != -1 ? hl : Hist{hl[0], hl[1], last};
                const Hist hr2 = hr[2] != -1 ? hr : Hist{hr[0], hr[1], last};
                ops.push_back({OP_SUM, kv.second.off, hidx.at(hl2) * K + hidx.at(hr2), 0, 0, 0, len, 1});
                continue;
            }
            for (const Succ &sl : finals(hl, last))
                for (const Succ &sr : finals(hr, last)) {
                    const int xy = first_id(sl.x ? sl.x : 3) * 3 + first_id(sr.x ? sr.x : 3);
                    ops.push_back({OP_DOT, kv.second.off, hidx.at(sl.h) * K + hidx.at(sr.h), xy,
                                   local_list(xy, kv.second.cl, kv.second.cr), 0, len, 1});
                }
        }
        end_stage(ob, false, true);
    }
    // every entry of J is written exactly once
    std::vector<int> seen((size_t)K * K, 0);
    const PlanStage &fs = stages.back();
    for (int o = fs.op_begin; o < fs.op_end; ++o) seen[ops[o].dst] += 1;
    for (int v : seen)
        if (v > 1) throw std::runtime_error("joint matrix entry written more than once");
}

}  // namespace itr
