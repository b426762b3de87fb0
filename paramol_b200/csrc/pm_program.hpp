// pm_program.hpp -- host-side builder of the topology "program" the E+F kernel executes.
//
// The kernel gives every conformation a TEAM of L = 32/CPW lanes (CPW conformations per warp).  All lanes of a warp run
// the same step loop; in step s lane l of every team executes record table[s][l].  The builder guarantees that the
// records of one step touch DISJOINT atoms, so the read-modify-write of the shared force tile needs no atomics
// (lanes of a team run in lockstep; a __syncwarp separates steps).  With L = 1 the tables degenerate to plain lists and
// every index / parameter read is a warp-uniform broadcast.
//
// Record kinds
//   star   : a centre atom with up to 4 neighbours: the bonds centre-neighbour and the angles neighbour-centre-neighbour
//            assigned to it.  Geometry (d_k = x_c - x_k, |d_k|^2, 1/|d_k|) is computed once per neighbour; angle forces
//            are linear combinations of the d_k (d_a x (d_a x d_b) = d_a (d_a.d_b) - d_b |d_a|^2), so no cross product is
//            formed, and the centre force is minus the sum of the neighbour forces.
//   axis   : a rotation axis j-k with up to 3 substituents i on j and up to 3 substituents l on k and every torsion
//            i-j-k-l on it (any number of periodicities per (i,l) cell).  c1_i = d0_i x d1 and c2_l = d1 x d2_l are
//            computed once per substituent; per torsion only scalars remain because every force is a linear combination
//            of c1_i and c2_l.  Impropers are axes with one cell.
//   pair segment : an i-tile of up to 4 atoms held in registers by one lane plus a list of (j, mask, pair-base) entries.
//            Tiles are cut into chunks so that the L lanes of a team get equal work; inside a segment the entries are
//            ordered so that no two lanes of a team update the same j in the same step.  (One lane per conformation only.)
//   pair block   : teams of L >= 4 lanes run the pair phase as 4 x 4 register blocks instead: atoms are dealt into tiles of 4
//            (tile t = atoms t, t + T, t + 2T, t + 3T, so that the rows the lanes of a team touch spread over the shared-memory
//            banks), a block is a pair of tiles (ti <= tj) with its 16 pair-parameter indices; pairs that do not interact
//            (excluded, padding atoms, the lower triangle of a diagonal block) point at a zero parameter slot, so every lane runs
//            the SAME branch-free 16-pair code on its own block -- no masks, no divergence between the lanes of a team.  A
//            position holds up to L blocks on pairwise disjoint tiles (both tiles' forces are flushed after the block, so no
//            two lanes of a team ever touch the same atom in one position).
#pragma once
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <map>
#include <set>
#include <vector>

#define PMH_IT 4

struct PmHostProgram {
    int n_atoms = 0, L = 1, cpw = 32;
    int team_warps = 0;         // > 1: teams of warps (see pm_build_program)
    std::vector<int> blob;      // everything the kernel stages into shared memory
    std::vector<int> pair_map;  // (i, j, exception or -1) per interacting pair, kernel order
    int n_pairs = 0;
    int n_star_steps = 0, n_axis_steps = 0, n_segs = 0;
    int n_pair_pos = 0, off_blocks = 0;  // team mode: 4 x 4 pair blocks (see below); then n_segs == 0
    int off_star_tab = 0, off_star_hdr = 0, off_star_rec = 0, off_axis_tab = 0, off_axis_hdr = 0, off_axis_rec = 0, off_axis_terms = 0;
    int off_seg_hdr = 0, off_seg_tile = 0, off_entries = 0, off_tper = 0;
    // statistics for DESIGN.md / pm_launch_info
    int n_stars = 0, n_axes = 0, n_entries = 0, n_entry_slots = 0;
};

namespace pmh {

struct Item {
    std::vector<int> atoms;
    int cost = 0;
    int rec = 0;
};

// Steps of up to L records.  Table entry (step, lane) = (record or -1, colour).  Records of one step that share a
// colour touch disjoint atoms; the kernel computes all records of a step in parallel and serialises only the force
// write-back by colour.  hdr[step] = number of colours.  Records are taken in order of decreasing cost (balance), each
// slot preferring, among the next few candidates, the one that conflicts least with the records already in the step.
inline std::vector<int> schedule(const std::vector<Item> &items, int L, int *n_steps, std::vector<int> &hdr, bool strict = false) {
    std::vector<int> table;
    hdr.clear();
    const int n = (int)items.size();
    if (L == 1) {
        for (int i = 0; i < n; ++i) {
            table.push_back(items[i].rec);
            table.push_back(0);
            hdr.push_back(1);
        }
        *n_steps = n;
        return table;
    }
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return items[a].cost > items[b].cost; });
    std::vector<char> done(n, 0);
    int left = n, steps = 0;
    auto overlaps = [&](int a, int b) {
        for (int x : items[a].atoms)
            for (int y : items[b].atoms)
                if (x == y) return true;
        return false;
    };
    while (left > 0) {
        std::vector<int> chosen;
        for (int slot = 0; slot < L && left > 0; ++slot) {
            int best = -1, best_conf = INT_MAX, seen = 0;
            // strict (teams of WARPS: the records of a step are written back concurrently, with one barrier per step): only
            // records that touch none of the atoms already in the step, searched over everything that is left
            for (int oi = 0; oi < n && (strict || seen < 2 * L); ++oi) {
                const int i = order[oi];
                if (done[i]) continue;
                ++seen;
                int conf = 0;
                for (int c : chosen) conf += overlaps(i, c);
                if (conf < best_conf) {
                    best_conf = conf;
                    best = i;
                }
                if (conf == 0) break;
            }
            if (strict && best_conf > 0) break;  // nothing disjoint is left: close the step
            chosen.push_back(best);
            done[best] = 1;
            --left;
        }
        std::vector<int> colour(chosen.size(), 0);
        int n_col = 1;
        for (size_t a = 0; a < chosen.size(); ++a) {
            std::set<int> used;
            for (size_t b = 0; b < a; ++b)
                if (overlaps(chosen[a], chosen[b])) used.insert(colour[b]);
            int c = 0;
            while (used.count(c)) ++c;
            colour[a] = c;
            n_col = std::max(n_col, c + 1);
        }
        for (int l = 0; l < L; ++l) {
            table.push_back(l < (int)chosen.size() ? items[chosen[l]].rec : -1);
            table.push_back(l < (int)chosen.size() ? colour[l] : 0);
        }
        hdr.push_back(n_col);
        ++steps;
    }
    *n_steps = steps;
    return table;
}

}  // namespace pmh

// cls[i*N+j]: 0 plain pair, 1 excluded, 2+e active exception e
inline bool pm_build_program(int N, int cpw, int n_bonds, const int *bond_idx, int n_angles, const int *angle_idx, int n_torsions,
                             const int *tors_idx, const int *tors_per, const std::vector<int> &cls, PmHostProgram &out, int team_warps = 0) {
    out = PmHostProgram();
    // team_warps > 1: a team is that many WARPS sharing one tile of 32 conformations (lane = conformation, cpw = 32); the tables are
    // those of a team of L members, but the records of a step must be disjoint (one colour)
    const bool warp_team = team_warps > 1;
    const int L = warp_team ? team_warps : 32 / cpw;
    out.team_warps = warp_team ? team_warps : 0;
    out.n_atoms = N;
    out.L = L;
    out.cpw = cpw;
    // atoms are stored as BYTE offsets of their first row inside a [3N][cpw] tile of doubles
    const int AO = 3 * cpw * 8;

    // ------------------------------------------------------------------ stars -------------------------------------
    // neighbour sets from the angle list (centre = middle atom), chunked to 4
    std::vector<std::vector<int>> nbr(N);
    auto add_nbr = [&](int c, int a) {
        if (std::find(nbr[c].begin(), nbr[c].end(), a) == nbr[c].end()) nbr[c].push_back(a);
    };
    for (int a = 0; a < n_angles; ++a) {
        add_nbr(angle_idx[3 * a + 1], angle_idx[3 * a]);
        add_nbr(angle_idx[3 * a + 1], angle_idx[3 * a + 2]);
    }
    struct Star {
        int c;
        std::vector<int> n;
        int bond[4], angle[6];
    };
    std::vector<Star> stars;
    auto new_star = [&](int c) {
        Star s;
        s.c = c;
        for (int k = 0; k < 4; ++k) s.bond[k] = -1;
        for (int k = 0; k < 6; ++k) s.angle[k] = -1;
        stars.push_back(s);
        return (int)stars.size() - 1;
    };
    static const int PAIR_A[6] = {0, 0, 0, 1, 1, 2}, PAIR_B[6] = {1, 2, 3, 2, 3, 3};
    auto pair_slot = [&](int a, int b) {
        if (a > b) std::swap(a, b);
        for (int q = 0; q < 6; ++q)
            if (PAIR_A[q] == a && PAIR_B[q] == b) return q;
        return -1;
    };
    std::vector<std::vector<int>> stars_of(N);
    for (int c = 0; c < N; ++c) {
        for (size_t k0 = 0; k0 < nbr[c].size(); k0 += 4) {
            const int si = new_star(c);
            for (size_t k = k0; k < std::min(nbr[c].size(), k0 + 4); ++k) stars[si].n.push_back(nbr[c][k]);
            stars_of[c].push_back(si);
        }
    }
    auto pos_in = [&](const Star &s, int a) {
        for (size_t k = 0; k < s.n.size(); ++k)
            if (s.n[k] == a) return (int)k;
        return -1;
    };
    for (int a = 0; a < n_angles; ++a) {
        const int i = angle_idx[3 * a], c = angle_idx[3 * a + 1], k = angle_idx[3 * a + 2];
        if (i == k || i == c || k == c) return false;  // degenerate angle
        bool placed = false;
        for (int si : stars_of[c]) {
            const int pi = pos_in(stars[si], i), pk = pos_in(stars[si], k);
            if (pi >= 0 && pk >= 0 && pi != pk) {
                const int q = pair_slot(pi, pk);
                if (stars[si].angle[q] < 0) {
                    stars[si].angle[q] = a;
                    placed = true;
                    break;
                }
            }
        }
        if (!placed) {  // neighbours in different chunks or a duplicated angle: a star of its own
            const int si = new_star(c);
            stars[si].n.push_back(i);
            stars[si].n.push_back(k);
            stars[si].angle[0] = a;
            stars_of[c].push_back(si);
        }
    }
    for (int b = 0; b < n_bonds; ++b) {
        const int i = bond_idx[2 * b], j = bond_idx[2 * b + 1];
        if (i == j) return false;
        bool placed = false;
        for (int pass = 0; pass < 2 && !placed; ++pass) {
            const int c = pass ? j : i, o = pass ? i : j;
            for (int si : stars_of[c]) {
                const int p = pos_in(stars[si], o);
                if (p >= 0 && stars[si].bond[p] < 0) {
                    stars[si].bond[p] = b;
                    placed = true;
                    break;
                }
            }
        }
        if (!placed) {
            // try to append the partner to a star of i or j that still has room, else a new star
            for (int pass = 0; pass < 2 && !placed; ++pass) {
                const int c = pass ? j : i, o = pass ? i : j;
                for (int si : stars_of[c])
                    if (stars[si].n.size() < 4 && pos_in(stars[si], o) < 0) {
                        stars[si].n.push_back(o);
                        stars[si].bond[stars[si].n.size() - 1] = b;
                        placed = true;
                        break;
                    }
            }
            if (!placed) {
                const int si = new_star(i);
                stars[si].n.push_back(j);
                stars[si].bond[0] = b;
                stars_of[i].push_back(si);
            }
        }
    }
    std::vector<pmh::Item> star_items;
    std::vector<int> star_rec;
    for (size_t s = 0; s < stars.size(); ++s) {
        const Star &st = stars[s];
        int nb = 0, na = 0;
        for (int k = 0; k < 4; ++k) nb += st.bond[k] >= 0;
        for (int k = 0; k < 6; ++k) na += st.angle[k] >= 0;
        if (nb + na == 0) continue;
        pmh::Item it;
        it.atoms.push_back(st.c);
        for (int a : st.n) it.atoms.push_back(a);
        it.cost = 13 * (int)st.n.size() + 8 * nb + 45 * na;
        it.rec = (int)star_items.size();
        star_items.push_back(it);
        star_rec.push_back(st.c * AO);
        star_rec.push_back((int)st.n.size());
        for (int k = 0; k < 4; ++k) star_rec.push_back((k < (int)st.n.size() ? st.n[k] : st.c) * AO);
        for (int k = 0; k < 4; ++k) star_rec.push_back(st.bond[k]);
        for (int k = 0; k < 6; ++k) star_rec.push_back(st.angle[k]);
    }
    out.n_stars = (int)star_items.size();

    // ------------------------------------------------------------------ axes --------------------------------------
    struct Axis {
        int j, k;
        std::vector<int> I, Lst;
        std::vector<std::vector<int>> cell;  // 9 lists of torsion indices
    };
    std::vector<Axis> axes;
    std::map<std::pair<int, int>, std::vector<int>> axes_of;  // (j,k) -> axis indices
    for (int t = 0; t < n_torsions; ++t) {
        int i = tors_idx[4 * t], j = tors_idx[4 * t + 1], k = tors_idx[4 * t + 2], l = tors_idx[4 * t + 3];
        if (j == k) return false;
        // canonical direction j < k: phi(i,j,k,l) == phi(l,k,j,i)
        if (j > k) {
            std::swap(i, l);
            std::swap(j, k);
        }
        bool placed = false;
        auto &lst = axes_of[std::make_pair(j, k)];
        for (int ai : lst) {
            Axis &ax = axes[ai];
            int pi = -1, pl = -1;
            for (size_t q = 0; q < ax.I.size(); ++q)
                if (ax.I[q] == i) pi = (int)q;
            for (size_t q = 0; q < ax.Lst.size(); ++q)
                if (ax.Lst[q] == l) pl = (int)q;
            if (pi < 0 && ax.I.size() >= 3) continue;
            if (pl < 0 && ax.Lst.size() >= 3) continue;
            if (pi < 0) {
                ax.I.push_back(i);
                pi = (int)ax.I.size() - 1;
            }
            if (pl < 0) {
                ax.Lst.push_back(l);
                pl = (int)ax.Lst.size() - 1;
            }
            ax.cell[pi * 3 + pl].push_back(t);
            placed = true;
            break;
        }
        if (!placed) {
            Axis ax;
            ax.j = j;
            ax.k = k;
            ax.I.push_back(i);
            ax.Lst.push_back(l);
            ax.cell.assign(9, std::vector<int>());
            ax.cell[0].push_back(t);
            axes.push_back(ax);
            lst.push_back((int)axes.size() - 1);
        }
    }
    std::vector<pmh::Item> axis_items;
    std::vector<int> axis_rec, axis_terms;
    for (size_t a = 0; a < axes.size(); ++a) {
        const Axis &ax = axes[a];
        pmh::Item it;
        it.atoms.push_back(ax.j);
        it.atoms.push_back(ax.k);
        for (int x : ax.I) it.atoms.push_back(x);
        for (int x : ax.Lst) it.atoms.push_back(x);
        // an atom may not appear twice in a record's RMW set (e.g. i == l in a 3-ring): handled because RMWs inside a
        // record are sequential per lane; only cross-lane disjointness matters
        std::sort(it.atoms.begin(), it.atoms.end());
        it.atoms.erase(std::unique(it.atoms.begin(), it.atoms.end()), it.atoms.end());
        int nterms = 0;
        for (auto &c : ax.cell) nterms += (int)c.size();
        it.cost = 14 + 34 * (int)(ax.I.size() + ax.Lst.size()) + 30 * nterms;
        it.rec = (int)axis_items.size();
        axis_items.push_back(it);
        axis_rec.push_back(ax.j * AO);
        axis_rec.push_back(ax.k * AO);
        // "simple" axis: every (i, l) cell holds exactly one torsion of periodicity 1..4 -> the kernel runs the cells of an
        // i-substituent as one branch-free block (independent dependency chains the compiler can interleave)
        bool simple = true;
        for (int pi = 0; pi < (int)ax.I.size(); ++pi)
            for (int pl = 0; pl < (int)ax.Lst.size(); ++pl) {
                const auto &c = ax.cell[pi * 3 + pl];
                if (c.size() != 1 || tors_per[c[0]] < 1 || tors_per[c[0]] > 4) simple = false;
            }
        axis_rec.push_back((int)ax.I.size() | (simple ? 256 : 0));
        axis_rec.push_back((int)ax.Lst.size());
        for (int q = 0; q < 3; ++q) axis_rec.push_back((q < (int)ax.I.size() ? ax.I[q] : ax.j) * AO);
        for (int q = 0; q < 3; ++q) axis_rec.push_back((q < (int)ax.Lst.size() ? ax.Lst[q] : ax.k) * AO);
        for (int q = 0; q < 9; ++q) {
            axis_rec.push_back((int)axis_terms.size() / 2);
            for (int t : ax.cell[q]) {
                axis_terms.push_back(t);
                axis_terms.push_back(tors_per[t]);
            }
        }
        axis_rec.push_back((int)axis_terms.size() / 2);
    }
    out.n_axes = (int)axis_items.size();

    // ------------------------------------------------------------------ pairs -------------------------------------
    struct Entry {
        int j, mask, base;
    };
    struct Unit {
        int tile;
        std::vector<Entry> e;
    };
    const int n_itiles = (N + PMH_IT - 1) / PMH_IT;
    std::vector<std::vector<Entry>> tile_entries(n_itiles);
    int total_entries = 0;
    for (int it = 0; it < n_itiles; ++it) {
        const int a0 = it * PMH_IT;
        for (int j = a0 + 1; j < N; ++j) {
            int mask = 0;
            const int base = (int)(out.pair_map.size() / 3);
            for (int k = 0; k < PMH_IT; ++k) {
                const int a = a0 + k;
                if (a >= N || a >= j) continue;
                const int c = cls[(size_t)a * N + j];
                if (c == 1) continue;
                mask |= 1 << k;
                out.pair_map.push_back(a);
                out.pair_map.push_back(j);
                out.pair_map.push_back(c == 0 ? -1 : c - 2);
            }
            if (mask) {
                tile_entries[it].push_back(Entry{j, mask, base});
                ++total_entries;
            }
        }
    }
    out.n_pairs = (int)(out.pair_map.size() / 3);
    out.n_entries = total_entries;

    // ---- team mode: 4 x 4 blocks on disjoint tile pairs ---------------------------------------------------------------
    std::vector<int> block_tab;
    const char *pb_env = getenv("PM_PAIR_BLOCKS");
    // (measured on B200: with 2 lanes per conformation the masked segments are still ahead -- caffeine 385 vs 362 M conformations/s;
    // from 4 lanes on the blocks win -- norfloxacin 142 -> 150, lig60 63 -> 71)
    const bool use_blocks = L >= 4 && out.n_pairs > 0 && out.n_pairs < 65535 && !(pb_env && pb_env[0] == '0');
    if (use_blocks) {
        // (measured and dropped for teams of warps: 2 L - 1 = 15 tiles of 3-4 atoms for 45-56 atoms, so that every position of the
        // tournament is full -- lig50 131.4 vs 134.9 M conformations/s with its 13 tiles of 4: the padding columns cost what the
        // idle member saves)
        const int T = (N + PMH_IT - 1) / PMH_IT;
        auto tile_atom = [&](int t, int k) {
            const int a = t + k * T;
            return a < N ? a : -1;
        };
        auto tile_size = [&](int t) {
            int n = 0;
            for (int k = 0; k < 4; ++k) n += tile_atom(t, k) >= 0;
            return n;
        };
        std::map<std::pair<int, int>, int> pair_index;
        for (int q = 0; q < out.n_pairs; ++q) pair_index[std::make_pair(out.pair_map[3 * q], out.pair_map[3 * q + 1])] = q;
        struct Block {
            int ti, tj;   // row tile, column tile
            int idx[16];
        };
        std::vector<Block> blocks;
        for (int t0 = 0; t0 < T; ++t0)
            for (int t1 = t0; t1 < T; ++t1) {
                // rows = the tile with fewer atoms (teams of warps skip padding rows)
                const bool flip = warp_team && tile_size(t1) < tile_size(t0);
                const int ti = flip ? t1 : t0, tj = flip ? t0 : t1;
                Block b;
                b.ti = ti;
                b.tj = tj;
                int real = 0;
                for (int k = 0; k < 4; ++k)
                    for (int m = 0; m < 4; ++m) {
                        int q = out.n_pairs;  // the zero slot
                        const int a = tile_atom(ti, k), c = tile_atom(tj, m);
                        if (a >= 0 && c >= 0 && a != c && (ti != tj || k < m)) {
                            auto it = pair_index.find(std::make_pair(std::min(a, c), std::max(a, c)));
                            if (it != pair_index.end()) {
                                q = it->second;
                                ++real;
                            }
                        }
                        b.idx[4 * k + m] = q;
                    }
                if (real) blocks.push_back(b);
            }
        // positions: greedy packing of up to L tile-disjoint blocks, busiest tiles first; a few deterministic perturbations of
        // the priorities, the shortest schedule wins
        std::vector<std::vector<int>> best;
        for (int attempt = 0; attempt < 64; ++attempt) {
            unsigned rng = 12345u + 7919u * (unsigned)attempt;
            std::vector<int> deg(T, 0);
            std::vector<char> done(blocks.size(), 0);
            for (const Block &b : blocks) {
                deg[b.ti]++;
                if (b.tj != b.ti) deg[b.tj]++;
            }
            size_t left = blocks.size();
            std::vector<std::vector<int>> sched;
            while (left > 0) {
                std::vector<std::pair<int, int>> cand;  // (-priority, block)
                for (size_t i = 0; i < blocks.size(); ++i)
                    if (!done[i]) {
                        rng = rng * 1664525u + 1013904223u;
                        const int jitter = attempt ? (int)((rng >> 24) & 3) : 0;
                        cand.push_back(std::make_pair(-(4 * (deg[blocks[i].ti] + deg[blocks[i].tj]) + jitter), (int)i));
                    }
                std::sort(cand.begin(), cand.end());
                std::vector<int> cur;
                std::set<int> used;
                for (auto &c : cand) {
                    if ((int)cur.size() >= L) break;
                    const Block &b = blocks[c.second];
                    if (used.count(b.ti) || used.count(b.tj)) continue;
                    cur.push_back(c.second);
                    used.insert(b.ti);
                    used.insert(b.tj);
                }
                for (int i : cur) {
                    done[i] = 1;
                    --left;
                    deg[blocks[i].ti]--;
                    if (blocks[i].tj != blocks[i].ti) deg[blocks[i].tj]--;
                }
                sched.push_back(cur);
            }
            if (best.empty() || sched.size() < best.size()) best = sched;
        }
        // teams of warps, odd tile count with (T + 1) / 2 <= L: the round-robin tournament (circle method) -- round r pairs the tiles
        // r + i and r - i (mod T), i = 1 .. (T - 1) / 2, and gives tile r its diagonal block: T positions, every tile exactly once
        // per position (lig60: 15 positions instead of the 17 the greedy packing finds)
        if (warp_team && (T & 1) && (T + 1) / 2 <= L) {
            std::map<std::pair<int, int>, int> block_of;
            for (size_t i = 0; i < blocks.size(); ++i)
                block_of[std::make_pair(std::min(blocks[i].ti, blocks[i].tj), std::max(blocks[i].ti, blocks[i].tj))] = (int)i;
            std::vector<std::vector<int>> rr;
            for (int r = 0; r < T; ++r) {
                std::vector<int> cur;
                for (int i = 0; i <= (T - 1) / 2; ++i) {
                    const int a = (r + i) % T, c = ((r - i) % T + T) % T;
                    auto it = block_of.find(std::make_pair(std::min(a, c), std::max(a, c)));
                    if (it != block_of.end()) cur.push_back(it->second);
                }
                if (!cur.empty()) rr.push_back(cur);
            }
            if (rr.size() < best.size()) best = rr;
        }
        const int dummy = N * AO;  // rows of the padding atom behind the tile (zero coordinates, never-read forces)
        for (auto &pos : best)
            for (int l = 0; l < L; ++l) {
                int rec[16];
                for (int k = 0; k < 8; ++k) rec[k] = dummy;
                for (int k = 0; k < 8; ++k) rec[8 + k] = out.n_pairs | (out.n_pairs << 16);
                if (l < (int)pos.size()) {
                    const Block &b = blocks[pos[l]];
                    for (int k = 0; k < 4; ++k) {
                        const int a = tile_atom(b.ti, k), c = tile_atom(b.tj, k);
                        rec[k] = a < 0 ? dummy : a * AO;
                        rec[4 + k] = c < 0 ? dummy : c * AO;
                    }
                    for (int k = 0; k < 8; ++k) rec[8 + k] = b.idx[2 * k] | (b.idx[2 * k + 1] << 16);
                }
                block_tab.insert(block_tab.end(), rec, rec + 16);
            }
        out.n_pair_pos = (int)best.size();
        out.n_entries = (int)blocks.size();
    }

    auto build_units = [&](int chunk) {
        std::vector<Unit> units;
        for (int it = 0; it < n_itiles; ++it) {
            const auto &te = tile_entries[it];
            if (te.empty()) continue;
            const int n_chunks = chunk >= (int)te.size() ? 1 : ((int)te.size() + chunk - 1) / chunk;
            const int per = ((int)te.size() + n_chunks - 1) / n_chunks;  // even split
            for (size_t o = 0; o < te.size(); o += per) {
                Unit u;
                u.tile = it;
                u.e.assign(te.begin() + o, te.begin() + std::min(te.size(), o + (size_t)per));
                units.push_back(u);
            }
        }
        return units;
    };
    // segments: groups of <= L units from distinct tiles, similar lengths
    auto build_segments = [&](const std::vector<Unit> &units, std::vector<std::vector<int>> &segs) {
        segs.clear();
        std::vector<int> order(units.size());
        for (size_t i = 0; i < units.size(); ++i) order[i] = (int)i;
        if (L > 1)
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return units[a].e.size() > units[b].e.size(); });
        std::vector<char> done(units.size(), 0);
        size_t left = units.size();
        long cost = 0;
        while (left > 0) {
            std::vector<int> seg;
            std::set<int> tiles;
            for (size_t oi = 0; oi < order.size() && (int)seg.size() < L; ++oi) {
                const int u = order[oi];
                if (done[u] || tiles.count(units[u].tile)) continue;
                seg.push_back(u);
                tiles.insert(units[u].tile);
                done[u] = 1;
                --left;
            }
            cost += (long)units[seg[0]].e.size() + 6;
            segs.push_back(seg);
        }
        return cost;
    };
    std::vector<Unit> units;
    std::vector<std::vector<int>> segs;
    if (use_blocks) {
        // pair blocks replace the segments
    } else if (L == 1 || total_entries == 0) {
        units = build_units(INT_MAX);
        build_segments(units, segs);
    } else {
        long best = LONG_MAX;
        int best_chunk = INT_MAX;
        for (int R = 1; R <= 24; ++R) {
            const int chunk = std::max(2, (total_entries + L * R - 1) / (L * R));
            std::vector<Unit> u = build_units(chunk);
            std::vector<std::vector<int>> s;
            const long c = build_segments(u, s);
            if (c < best) {
                best = c;
                best_chunk = chunk;
            }
        }
        units = build_units(best_chunk);
        build_segments(units, segs);
    }
    // emit: per segment, order entries so that no two lanes touch the same j at the same position
    std::vector<int> seg_hdr, seg_tile, entries;
    int slots = 0;
    for (auto &seg : segs) {
        std::vector<std::vector<Entry>> lists(L);
        for (size_t l = 0; l < seg.size(); ++l) lists[l] = units[seg[l]].e;
        seg_hdr.push_back(0);  // n_e, patched below
        seg_hdr.push_back((int)entries.size() / 2);
        for (int l = 0; l < L; ++l)
            for (int k = 0; k < PMH_IT; ++k) {
                int a = -1;
                if (l < (int)seg.size()) {
                    a = units[seg[l]].tile * PMH_IT + k;
                    if (a >= N) a = -1;
                }
                seg_tile.push_back(a < 0 ? -1 : a * AO);
            }
        int n_e = 0;
        while (true) {
            bool any = false;
            for (int l = 0; l < L; ++l) any = any || !lists[l].empty();
            if (!any) break;
            std::set<int> used;
            for (int l = 0; l < L; ++l) {
                int pick = -1;
                for (size_t q = 0; q < lists[l].size(); ++q)
                    if (!used.count(lists[l][q].j)) {
                        pick = (int)q;
                        break;
                    }
                if (pick < 0) {
                    entries.push_back(0);  // bubble: mask 0
                    entries.push_back(0);
                } else {
                    const Entry e = lists[l][pick];
                    lists[l].erase(lists[l].begin() + pick);
                    used.insert(e.j);
                    entries.push_back((e.j * AO) | (e.mask << 28));
                    entries.push_back(e.base);
                }
                ++slots;
            }
            ++n_e;
        }
        seg_hdr[seg_hdr.size() - 2] = n_e;
    }
    out.n_segs = (int)segs.size();
    out.n_entry_slots = use_blocks ? out.n_pair_pos * L : slots;

    // ------------------------------------------------------------------ tables + blob ---------------------------------
    std::vector<int> star_hdr, axis_hdr;
    // teams of warps: steps of records in several colours, one round of the team's split barrier per colour (default), or
    // PM_WT_COLOURS=0: strictly disjoint steps of one colour (fewer rounds, but few records find room in a step: lig50's 27 axes
    // take 8 steps of 8 members instead of 4)
    const char *wt_col = std::getenv("PM_WT_COLOURS");
    const bool strict = warp_team && wt_col != nullptr && wt_col[0] == '0';
    std::vector<int> star_tab = pmh::schedule(star_items, L, &out.n_star_steps, star_hdr, strict);
    std::vector<int> axis_tab = pmh::schedule(axis_items, L, &out.n_axis_steps, axis_hdr, strict);
    // axis step header: colours | (largest number of i-substituents in the step << 8)
    for (int st = 0; st < out.n_axis_steps; ++st) {
        int max_ni = 0;
        for (int l = 0; l < L; ++l) {
            const int rec = axis_tab[2 * (st * L + l)];
            if (rec >= 0) max_ni = std::max(max_ni, axis_rec[20 * rec + 2] & 255);
        }
        axis_hdr[st] |= max_ni << 8;
    }
    auto align4 = [&]() {
        while (out.blob.size() & 3) out.blob.push_back(0);
    };
    auto append = [&](const std::vector<int> &v) {
        align4();
        const int o = (int)out.blob.size();
        out.blob.insert(out.blob.end(), v.begin(), v.end());
        return o;
    };
    out.off_star_tab = append(star_tab);
    out.off_star_hdr = append(star_hdr);
    out.off_star_rec = append(star_rec);
    out.off_axis_tab = append(axis_tab);
    out.off_axis_hdr = append(axis_hdr);
    out.off_axis_rec = append(axis_rec);
    out.off_axis_terms = append(axis_terms);
    out.off_seg_hdr = append(seg_hdr);
    out.off_seg_tile = append(seg_tile);
    out.off_entries = append(entries);
    out.off_tper = append(std::vector<int>(tors_per, tors_per + n_torsions));
    out.off_blocks = append(block_tab);
    align4();
    return true;
}

// ---------------------------------------------------------------------------------------------------------------------
// Host-only validator (used by the CPU test-suite through pm_program_check): every bond / angle / torsion / interacting
// pair is executed exactly once and records sharing a step touch disjoint atoms.  Returns 0 when all invariants hold.
// ---------------------------------------------------------------------------------------------------------------------
inline int pm_validate_program(const PmHostProgram &p, int N, int n_bonds, const int *bond_idx, int n_angles, const int *angle_idx,
                               int n_torsions, const int *tors_idx, const std::vector<int> &cls) {
    const int L = p.L;
    const int AO = 3 * p.cpw * 8;
    const int *b = p.blob.data();
    std::vector<int> seen_b(n_bonds, 0), seen_a(n_angles, 0), seen_t(n_torsions, 0);
    static const int PA[6] = {0, 0, 0, 1, 1, 2}, PB[6] = {1, 2, 3, 2, 3, 3};
    // stars
    for (int s = 0; s < p.n_star_steps; ++s) {
        std::map<int, std::set<int>> used_by_colour;
        const int n_col = b[p.off_star_hdr + s];
        for (int l = 0; l < L; ++l) {
            const int rec = b[p.off_star_tab + 2 * (s * L + l)], colour = b[p.off_star_tab + 2 * (s * L + l) + 1];
            if (rec < 0) continue;
            if (colour < 0 || colour >= n_col) return 17;
            std::set<int> &used = used_by_colour[colour];
            const int *r = b + p.off_star_rec + 16 * rec;
            const int c = r[0] / AO, m = r[1];
            if (m < 1 || m > 4 || r[0] % AO) return 10;
            std::set<int> mine;
            mine.insert(c);
            for (int k = 0; k < m; ++k)
                if (!mine.insert(r[2 + k] / AO).second) return 11;
            for (int a : mine)
                if (!used.insert(a).second) return 12;  // cross-lane conflict
            for (int k = 0; k < 4; ++k) {
                const int bi = r[6 + k];
                if (bi < 0) continue;
                if (k >= m || bi >= n_bonds) return 13;
                const int x = bond_idx[2 * bi], y = bond_idx[2 * bi + 1];
                if (!((x == c && y == r[2 + k] / AO) || (y == c && x == r[2 + k] / AO))) return 14;
                seen_b[bi]++;
            }
            for (int q = 0; q < 6; ++q) {
                const int ai = r[10 + q];
                if (ai < 0) continue;
                if (PB[q] >= m || ai >= n_angles) return 15;
                const int x = angle_idx[3 * ai], cc = angle_idx[3 * ai + 1], y = angle_idx[3 * ai + 2];
                const int na = r[2 + PA[q]] / AO, nb = r[2 + PB[q]] / AO;
                if (cc != c || !((x == na && y == nb) || (x == nb && y == na))) return 16;
                seen_a[ai]++;
            }
        }
    }
    // axes
    for (int s = 0; s < p.n_axis_steps; ++s) {
        std::map<int, std::set<int>> used_by_colour;
        const int n_col = b[p.off_axis_hdr + s] & 0xff, max_ni = b[p.off_axis_hdr + s] >> 8;
        for (int l = 0; l < L; ++l) {
            const int rec = b[p.off_axis_tab + 2 * (s * L + l)], colour = b[p.off_axis_tab + 2 * (s * L + l) + 1];
            if (rec < 0) continue;
            if (colour < 0 || colour >= n_col) return 27;
            std::set<int> &used = used_by_colour[colour];
            const int *r = b + p.off_axis_rec + 20 * rec;
            const int j = r[0] / AO, k = r[1] / AO, nI = r[2] & 255, nL = r[3];
            if (nI < 1 || nI > 3 || nL < 1 || nL > 3 || nI > max_ni) return 20;
            std::set<int> mine;
            mine.insert(j);
            mine.insert(k);
            for (int q = 0; q < nI; ++q) mine.insert(r[4 + q] / AO);
            for (int q = 0; q < nL; ++q) mine.insert(r[7 + q] / AO);
            for (int a : mine)
                if (!used.insert(a).second) return 21;
            for (int cell = 0; cell < 9; ++cell) {
                const int t0 = r[10 + cell], t1 = r[11 + cell];
                if (t1 < t0) return 22;
                const int pi = cell / 3, pl = cell % 3;
                if (t1 > t0 && (pi >= nI || pl >= nL)) return 23;
                for (int t = t0; t < t1; ++t) {
                    const int ti = b[p.off_axis_terms + 2 * t], per = b[p.off_axis_terms + 2 * t + 1];
                    if (ti < 0 || ti >= n_torsions) return 24;
                    const int *x = tors_idx + 4 * ti;
                    const bool fwd = x[0] == r[4 + pi] / AO && x[1] == j && x[2] == k && x[3] == r[7 + pl] / AO;
                    const bool rev = x[3] == r[4 + pi] / AO && x[2] == j && x[1] == k && x[0] == r[7 + pl] / AO;
                    if (!fwd && !rev) return 25;
                    if (per != b[p.off_tper + ti]) return 26;
                    seen_t[ti]++;
                }
            }
        }
    }
    for (int v : seen_b)
        if (v != 1) return 30;
    for (int v : seen_a)
        if (v != 1) return 31;
    for (int v : seen_t)
        if (v != 1) return 32;
    // pairs
    std::vector<int> seen_p(p.n_pairs, 0);
    for (int s = 0; s < p.n_segs; ++s) {
        const int n_e = b[p.off_seg_hdr + 2 * s], eoff = b[p.off_seg_hdr + 2 * s + 1];
        std::set<int> tiles;
        for (int l = 0; l < L; ++l) {
            const int a0 = b[p.off_seg_tile + (s * L + l) * 4];
            if (a0 >= 0 && !tiles.insert(a0).second) return 40;  // two lanes flushing the same tile
        }
        for (int e = 0; e < n_e; ++e) {
            std::set<int> js;
            for (int l = 0; l < L; ++l) {
                const int w0 = b[p.off_entries + 2 * (eoff + e * L + l)], base = b[p.off_entries + 2 * (eoff + e * L + l) + 1];
                const int mask = (w0 >> 28) & 15, j = (w0 & 0x0fffffff) / AO;
                if (!mask) continue;
                if (!js.insert(j).second) return 41;  // two lanes updating the same j in one step
                int q = base;
                for (int k = 0; k < PMH_IT; ++k) {
                    if (!(mask & (1 << k))) continue;
                    const int ao = b[p.off_seg_tile + (s * L + l) * 4 + k];
                    const int a = ao < 0 ? -1 : ao / AO;
                    if (a < 0 || q >= p.n_pairs) return 42;
                    if (p.pair_map[3 * q] != a || p.pair_map[3 * q + 1] != j) return 43;
                    seen_p[q]++;
                    ++q;
                }
            }
        }
    }
    // pair blocks (team mode): tiles of one position are disjoint across the lanes, every interacting pair appears once
    for (int pos = 0; pos < p.n_pair_pos; ++pos) {
        std::set<int> touched;
        for (int l = 0; l < L; ++l) {
            const int *r = b + p.off_blocks + 16 * (pos * L + l);
            std::set<int> mine;
            int at[8];
            for (int k = 0; k < 8; ++k) {
                if (r[k] % AO) return 50;
                at[k] = r[k] / AO;
                if (at[k] < 0 || at[k] > N) return 51;
                if (at[k] < N) mine.insert(at[k]);
            }
            for (int a : mine)
                if (!touched.insert(a).second) return 52;  // two lanes of a team on the same atom
            for (int k = 0; k < 4; ++k)
                for (int m = 0; m < 4; ++m) {
                    const int w = r[8 + (4 * k + m) / 2];
                    const int q = ((4 * k + m) & 1) ? (w >> 16) & 0xffff : w & 0xffff;
                    if (q == p.n_pairs) continue;  // zero slot
                    if (q > p.n_pairs) return 53;
                    const int a = at[k], c = at[4 + m];
                    if (a >= N || c >= N || a == c) return 54;
                    if (p.pair_map[3 * q] != std::min(a, c) || p.pair_map[3 * q + 1] != std::max(a, c)) return 55;
                    seen_p[q]++;
                }
        }
    }
    for (int v : seen_p)
        if (v != 1) return 44;
    // pair list == all non-excluded pairs
    int want = 0;
    for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j) want += cls[(size_t)i * N + j] != 1;
    if (want != p.n_pairs) return 45;
    for (int q = 0; q < p.n_pairs; ++q) {
        const int i = p.pair_map[3 * q], j = p.pair_map[3 * q + 1], e = p.pair_map[3 * q + 2];
        const int c = cls[(size_t)i * N + j];
        if (i >= j || c == 1 || (c == 0) != (e < 0) || (c >= 2 && e != c - 2)) return 46;
    }
    return 0;
}
