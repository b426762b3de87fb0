# -*- coding: utf-8 -*-
"""
Topology-specialised E+F kernel generator.

The generic kernel ``pm_ef_kernel`` (pm_kernels.cu) interprets a topology "program": every atom index, parameter offset
and loop bound is data.  For a given molecule all of that is known when the handle is created, so this module emits a CUDA
kernel in which the topology is CODE: straight-line blocks with immediate shared-memory offsets, the derived parameter
table in ``__constant__`` memory (parameters are direct constant-bank operands of the FP64 instructions), forces accumulated in registers (small molecules) or in the shared force tile with immediate offsets, and
one big basic block per conformation that lets ptxas interleave independent terms to cover the 8-cycle DFMA latency.
The arithmetic (trig-free torsions, table acos, c12/c6 pairs, the residual epilogue) is that of the generic kernel, term by
term, so results agree to rounding (measured: 5e-16 relative on E, F and the residual).

Same scope as the generic kernel for: one lane per conformation (CPW = 32), NoCutoff, full evaluation (mode 0).
"""
import numpy as np


class KernelSource:
    def __init__(self, n_atoms, bonds, angles, torsions, tors_per, pairs, f_in_registers=True, name="pm_ef_spec_kernel",
                 x_in_registers=False, fence_every=0, sync_every=0, with_forces=True):
        """
        bonds [nb,2], angles [na,3], torsions [nt,4], tors_per [nt], pairs [np,3] = (i, j, exception or -1) in the
        order of the derived table (pm_get_pairs).  Derived layout: bonds (r0,k) | angles (t0,k) | torsions
        (k, kn, cos, sin) | pairs (c12, c6, Kqq).
        """
        self.N = int(n_atoms)
        self.bonds = [tuple(int(v) for v in b) for b in np.asarray(bonds).reshape(-1, 2)]
        self.angles = [tuple(int(v) for v in a) for a in np.asarray(angles).reshape(-1, 3)]
        self.torsions = [tuple(int(v) for v in t) for t in np.asarray(torsions).reshape(-1, 4)]
        self.tors_per = [int(v) for v in np.asarray(tors_per).reshape(-1)]
        self.pairs = [tuple(int(v) for v in p) for p in np.asarray(pairs).reshape(-1, 3)]
        # forces in registers: True (all atoms), False (none: shared force tile) or an iterable of atom indices (the rest go to the tile)
        self.f_reg_atoms = set(range(int(n_atoms))) if f_in_registers is True else set(int(a) for a in (f_in_registers or ()))
        self.f_reg = len(self.f_reg_atoms) == int(n_atoms)
        # atoms whose forces live in the shared force tile, packed: tile row of atom a = 3 * f_tile_index[a]
        self.f_tile_atoms = [a for a in range(int(n_atoms)) if a not in self.f_reg_atoms]
        self.f_tile_index = {a: i for i, a in enumerate(self.f_tile_atoms)}
        self.x_reg = bool(x_in_registers)
        self.fence_every = int(fence_every)
        self.sync_every = int(sync_every)
        import os as _os
        self.pair_sync_every = int(_os.environ.get("PM_SPEC_PSYNC", PAIR_SYNC_EVERY))
        self.wf = bool(with_forces)
        # second shared tile per warp: the reference forces of the tile arrive by bulk copy while the body runs, and the next
        # tile's coordinates are requested as soon as the body has read the current ones (set by image_for when it fits)
        self.prefetch = 0
        # residual sums of pm_eval_reduce accumulated by the kernel itself (no [S] energy / residual round trip, no stage-1 launch)
        self.fused_sums = bool(int(_os.environ.get("PM_SPEC_FUSED", FUSED_SUMS)))
        self._blocks = 0
        self.name = name
        self.off_bond = 0
        self.off_angle = 2 * len(self.bonds)
        self.off_tors = self.off_angle + 2 * len(self.angles)
        self.off_pair = self.off_tors + 4 * len(self.torsions)
        self.n_derived = self.off_pair + 3 * len(self.pairs)
        self.lines = []

    # ---- emission helpers -----------------------------------------------------------------------
    def emit(self, text):
        self.lines.append(text)

    def P(self, off):
        return "spec_tab[%d]" % off

    def X(self, atom):
        return ("X%d" % atom) if self.x_reg else ("PMX(%d)" % atom)

    def end_block(self):
        """Close a star / axis / pair block; every ``fence_every`` blocks a warp-level fence keeps ptxas from hoisting the
        loads of all later blocks to the top (which costs more registers than the interleaving gains)."""
        self.emit("}")
        self._blocks += 1
        psync = self.pair_sync_every
        if psync and self._blocks % psync == 0 and not (self.sync_every and self._blocks % self.sync_every == 0):
            # the warps that share a scheduler (and its instruction buffer: warp, warp + 4, ...) wait for each other only
            self.emit('asm volatile("bar.sync %0, %1;" :: "r"(1 + (warp & 3)), "r"(32 * ((n_warps + 3 - (warp & 3)) / 4)) : "memory");')
            return
        if self.sync_every and self._blocks % self.sync_every == 0:
            # keeps the warps of the CTA within the same stretch of this (instruction-cache sized) straight-line code, so that a
            # fetched line serves all of them; the trip count of the tile loop is uniform per CTA for that reason
            self.emit("__syncthreads();")
        elif self.fence_every and self._blocks % self.fence_every == 0:
            # a fence ptxas honours (an empty asm with a memory clobber only constrains NVVM; without stores between them, as in
            # the energy-only kernel, ptxas still hoists every coordinate load to the top and spills)
            self.emit("__syncwarp();")

    def addf(self, atom, expr_vec, sign="+"):
        if not self.wf:
            return
        if atom in self.f_reg_atoms:
            self.emit("F%d.x %s= %s.x; F%d.y %s= %s.y; F%d.z %s= %s.z;" % (atom, sign, expr_vec, atom, sign, expr_vec, atom, sign, expr_vec))
        else:
            self.emit("PMF_%s(%d, %s);" % ("ADD" if sign == "+" else "SUB", self.f_tile_index[atom], expr_vec))

    # ---- stars: bonds + angles around a centre ---------------------------------------------------
    def gen_stars(self):
        N = self.N
        bond_of = {}
        for t, (i, j) in enumerate(self.bonds):
            bond_of[(i, j)] = t
            bond_of[(j, i)] = t
        angle_at = [[] for _ in range(N)]
        for t, (a, b, c) in enumerate(self.angles):
            angle_at[b].append((t, a, c))
        done_bond = set()
        # geometry neighbours of a centre: every atom it shares a bond or an angle with
        for c in range(N):
            need = []
            for (i, j) in self.bonds:
                if i == c and j not in need:
                    need.append(j)
                if j == c and i not in need:
                    need.append(i)
            for (_, a, b) in angle_at[c]:
                for v in (a, b):
                    if v not in need:
                        need.append(v)
            own = [n for n in need if (c, n) in bond_of and bond_of[(c, n)] not in done_bond]
            if not own and not angle_at[c]:
                continue
            used = set(own)
            for (_, a, b) in angle_at[c]:
                used.update((a, b))
            need = [n for n in need if n in used]
            self.emit("{  // star %d" % c)
            self.emit("const pm_d3 xc = %s;" % self.X(c))
            if self.wf:
                self.emit("pm_d3 fc = pm_d3{0.0, 0.0, 0.0};")
            for n in need:
                self.emit("const pm_d3 d%d = pm_sub(xc, %s); const double q%d = pm_dot(d%d, d%d); const double i%d = pm_rsqrt(q%d);"
                          % (n, self.X(n), n, n, n, n, n))
                if self.wf:
                    self.emit("pm_d3 f%d = pm_d3{0.0, 0.0, 0.0};" % n)
            for n in own:
                t = bond_of[(c, n)]
                done_bond.add(t)
                o = self.off_bond + 2 * t
                self.emit("{ const double dl = fma(q%d, i%d, -%s); const double kd = %s * dl; e_bond = fma(0.5 * kd, dl, e_bond);" % (n, n, self.P(o), self.P(o + 1))
                          + (" pm_axpy(f%d, kd * i%d, d%d); }" % (n, n, n) if self.wf else " }"))
            for (t, a, b) in angle_at[c]:
                o = self.off_angle + 2 * t
                # d_a = x_centre - x_a, d_b = x_centre - x_b (ReferenceAngleBondIxn).  As in the generic kernel: sin from the Lagrange
                # identity |d_a x d_b|^2 = |d_a|^2 |d_b|^2 - (d_a.d_b)^2 and the forces from the triple-product rule
                # d_a x (d_a x d_b) = d_a (d_a.d_b) - d_b |d_a|^2, so no cross product is formed (20 FP64 operations fewer per angle)
                self.emit("{ const double dt = pm_dot(d%d, d%d); const double i01 = i%d * i%d; const double pp2 = fmax(fma(-dt, dt, q%d * q%d), 0.0);" % (a, b, a, b, a, b))
                self.emit("  const double ip = pm_rsqrt(fmax(pp2, 1.0e-12)); const double th = pm_acos_cs(dt * i01, pp2 * ip * i01, s_acos);")
                self.emit("  const double dl = th - %s; const double dd = %s * dl; e_angle = fma(0.5 * dd, dl, e_angle);" % (self.P(o), self.P(o + 1)))
                if self.wf:
                    self.emit("  const double w = dd * ip, wd = w * dt; const double al = wd * (i%d * i%d), be = wd * (i%d * i%d);" % (a, a, b, b))
                    self.emit("  pm_axpy(f%d, al, d%d); pm_axpy(f%d, -w, d%d); pm_axpy(f%d, be, d%d); pm_axpy(f%d, -w, d%d); }" % (a, a, a, b, b, b, b, a))
                else:
                    self.emit("}")
            if self.wf and need:
                # the centre takes minus the sum of its neighbours' forces (once per star instead of once per term)
                for comp in ("x", "y", "z"):
                    self.emit("fc.%s = -(%s);" % (comp, " + ".join("f%d.%s" % (n, comp) for n in need)))
            self.addf(c, "fc")
            for n in need:
                self.addf(n, "f%d" % n)
            self.end_block()
        assert len(done_bond) == len(self.bonds)

    # ---- axes: all torsions around j-k -----------------------------------------------------------
    def gen_axes(self):
        axes = {}
        for t, (i, j, k, l) in enumerate(self.torsions):
            if (k, j) in axes:                      # same axis, opposite orientation: phi(l,k,j,i) = phi(i,j,k,l)
                axes[(k, j)].append((t, l, i))
            else:
                axes.setdefault((j, k), []).append((t, i, l))
        for (j, k), terms in axes.items():
            subs_i = sorted({i for (_, i, _) in terms})
            subs_l = sorted({l for (_, _, l) in terms})
            self.emit("{  // axis %d-%d" % (j, k))
            self.emit("const pm_d3 xj = %s, xk = %s; const pm_d3 d1 = pm_sub(xk, xj); const double bb = pm_dot(d1, d1);" % (self.X(j), self.X(k)))
            self.emit("const double rbb = pm_rsqrt(bb); const double nb = bb * rbb, ibb = rbb * rbb;")
            if self.wf:
                self.emit("pm_d3 fj = pm_d3{0.0, 0.0, 0.0}, fk = pm_d3{0.0, 0.0, 0.0};")
            # Forces of an axis: with A_i = -|d1| / |u_i|^2 * sum_l dE/dphi(i, l) and B_l = |d1| / |w_l|^2 * sum_i dE/dphi(i, l) the
            # substituents take F_i = A_i u_i, F_l = B_l w_l and the axis atoms F_j = sum_i (g_i - 1) F_i - sum_l h_l F_l,
            # F_k = sum_l (h_l - 1) F_l - sum_i g_i F_i (ReferenceProperDihedralBond), so a cell only adds its dE/dphi to the two
            # scalar sums of its substituents and the vector work is done once per substituent instead of once per cell
            for i in subs_i:
                self.emit("const pm_d3 a%d = pm_sub(%s, xj); const pm_d3 u%d = pm_cross(a%d, d1); const double m%d = pm_dot(u%d, u%d);"
                          " const double s%d = pm_rsqrt(m%d);" % (i, self.X(i), i, i, i, i, i, i, i)
                          + (" double za%d = 0.0;" % i if self.wf else ""))
            for l in subs_l:
                self.emit("const pm_d3 b%d = pm_sub(xk, %s); const pm_d3 w%d = pm_cross(d1, b%d); const double n%d = pm_dot(w%d, w%d);"
                          " const double t%d = pm_rsqrt(n%d);" % (l, self.X(l), l, l, l, l, l, l, l)
                          + (" double zb%d = 0.0;" % l if self.wf else ""))
            cells = {}
            for (t, i, l) in terms:
                cells.setdefault((i, l), []).append(t)
            for (i, l), ts in cells.items():
                ts = sorted(ts, key=lambda t: self.tors_per[t])
                self.emit("{ const double i12 = s%d * t%d; const double cphi = pm_dot(u%d, w%d) * i12, sphi = nb * pm_dot(a%d, w%d) * i12;" % (i, l, i, l, i, l))
                # (cos, sin)(n phi) by the angle-addition recurrence, started at n = 1 (n = 0 only if a term has periodicity 0)
                if self.tors_per[ts[0]] >= 1:
                    self.emit("  double cn = cphi, sn = sphi;")
                    cur = 1
                else:
                    self.emit("  double cn = 1.0, sn = 0.0;")
                    cur = 0
                if self.wf and len(ts) > 1:
                    self.emit("  double dd = 0.0;")
                for t in ts:
                    per = self.tors_per[t]
                    while cur < per:
                        self.emit("  { const double tt = cn * cphi - sn * sphi; sn = fma(sn, cphi, cn * sphi); cn = tt; }")
                        cur += 1
                    o = self.off_tors + 4 * t
                    self.emit("  e_tors += %s * (1.0 + fma(cn, %s, sn * %s));" % (self.P(o), self.P(o + 2), self.P(o + 3)))
                    if self.wf:
                        x = "fma(sn, %s, -(cn * %s))" % (self.P(o + 2), self.P(o + 3))   # sin(n phi - delta); dE/dphi = -k n sin(.)
                        if len(ts) > 1:
                            self.emit("  dd = fma(-%s, %s, dd);" % (self.P(o + 1), x))
                        else:
                            self.emit("  { const double x = %s; za%d = fma(-%s, x, za%d); zb%d = fma(-%s, x, zb%d); }"
                                      % (x, i, self.P(o + 1), i, l, self.P(o + 1), l))
                if self.wf and len(ts) > 1:
                    self.emit("  za%d += dd; zb%d += dd;" % (i, l))
                self.emit("}")
            if self.wf:
                for i in subs_i:
                    self.emit("const double g%d = pm_dot(a%d, d1) * ibb; const double A%d = -(nb * (s%d * s%d)) * za%d;"
                              " const pm_d3 fi%d = pm_d3{A%d * u%d.x, A%d * u%d.y, A%d * u%d.z};" % (i, i, i, i, i, i, i, i, i, i, i, i, i))
                    self.emit("pm_axpy(fj, g%d - 1.0, fi%d); pm_axpy(fk, -g%d, fi%d);" % (i, i, i, i))
                for l in subs_l:
                    self.emit("const double h%d = pm_dot(b%d, d1) * ibb; const double B%d = (nb * (t%d * t%d)) * zb%d;"
                              " const pm_d3 fl%d = pm_d3{B%d * w%d.x, B%d * w%d.y, B%d * w%d.z};" % (l, l, l, l, l, l, l, l, l, l, l, l, l))
                    self.emit("pm_axpy(fk, h%d - 1.0, fl%d); pm_axpy(fj, -h%d, fl%d);" % (l, l, l, l))
            self.addf(j, "fj")
            self.addf(k, "fk")
            for i in subs_i:
                self.addf(i, "fi%d" % i)
            for l in subs_l:
                self.addf(l, "fl%d" % l)
            self.end_block()

    # ---- pairs ---------------------------------------------------------------------------------
    def gen_pairs(self):
        by_i = {}
        for q, (i, j, _) in enumerate(self.pairs):
            by_i.setdefault(i, []).append((q, j))
        for i, lst in by_i.items():
            self.emit("{  // pairs of atom %d" % i)
            self.emit("const pm_d3 xi = %s;" % self.X(i) + (" pm_d3 fi = pm_d3{0.0, 0.0, 0.0};" if self.wf else ""))
            for (q, j) in lst:
                o = self.off_pair + 3 * q
                self.emit("{ const pm_d3 d = pm_sub(xi, %s); const double rinv = pm_rsqrt(pm_dot(d, d)); const double u = rinv * rinv; const double u3 = u * u * u;" % self.X(j))
                # b = c12 u^3 - c6 (one fma), E = b u^3 + K / r;  r dE/dr = -(6 u^3 (2 c12 u^3 - c6) + K / r), with 2 c12 u^3 - c6 = c12 u^3 + b
                self.emit("  const double b = fma(%s, u3, -%s); const double cq = %s * rinv; e_nb = fma(b, u3, e_nb); e_nb += cq;" % (self.P(o), self.P(o + 1), self.P(o + 2)))
                if self.wf:
                    self.emit("  const double g = fma(6.0 * u3, fma(%s, u3, b), cq) * u; const pm_d3 fv = pm_d3{g * d.x, g * d.y, g * d.z}; pm_axpy(fi, g, d);" % self.P(o))
                self.addf(j, "fv", "-")
                self.emit("}")
            self.addf(i, "fi")
            self.end_block()

    # ---- the kernel ----------------------------------------------------------------------------
    def source(self, warps, with_header=True):
        N = self.N
        self.lines = []
        self._blocks = 0
        self.gen_stars()
        self.gen_axes()
        self.gen_pairs()
        body = "\n            ".join(self.lines)
        f_decl = "\n            ".join("pm_d3 F%d = pm_d3{0.0, 0.0, 0.0};" % a for a in sorted(self.f_reg_atoms))
        if not self.f_reg:
            f_decl += "\n            for (int k = lane; k < %d; k += 32) fs[k] = 0.0; __syncwarp();" % (3 * len(self.f_tile_atoms) * 32)
        if not self.wf:
            f_decl = ""
        res_lines = []
        for a in range(N):
            if a in self.f_reg_atoms:
                fx, fy, fz = "F%d.x" % a, "F%d.y" % a, "F%d.z" % a
            else:
                ti = self.f_tile_index[a]
                fx, fy, fz = "fs[%d * 32 + lane]" % (3 * ti), "fs[%d * 32 + lane]" % (3 * ti + 1), "fs[%d * 32 + lane]" % (3 * ti + 2)
            res_lines.append("{ const double dx = %s - fr[%d * 32], dy = %s - fr[%d * 32], dz = %s - fr[%d * 32];" % (fx, 3 * a, fy, 3 * a + 1, fz, 3 * a + 2))
            m = ["spec_tab[SPEC_ND + %d]" % (9 * a + q) for q in range(9)]
            res_lines.append("  const double t0 = fma(dz, %s, fma(dy, %s, dx * %s)), t1 = fma(dz, %s, fma(dy, %s, dx * %s)), t2 = fma(dz, %s, fma(dy, %s, dx * %s));"
                             % (m[6], m[3], m[0], m[7], m[4], m[1], m[8], m[5], m[2]))
            res_lines.append("  res += fma(t2, dz, fma(t1, dy, t0 * dx)); }")
        res = "\n                ".join(res_lines) if self.wf else ""
        if not self.wf:
            fout = ""
        else:
            fout = "\n                ".join(
                ("fo[%d * 32] = F%d.x; fo[%d * 32] = F%d.y; fo[%d * 32] = F%d.z;" % (3 * a, a, 3 * a + 1, a, 3 * a + 2, a)) if a in self.f_reg_atoms else
                ("fo[%d * 32] = fs[%d * 32 + lane]; fo[%d * 32] = fs[%d * 32 + lane]; fo[%d * 32] = fs[%d * 32 + lane];"
                 % (3 * a, 3 * self.f_tile_index[a], 3 * a + 1, 3 * self.f_tile_index[a] + 1, 3 * a + 2, 3 * self.f_tile_index[a] + 2))
                for a in range(N))
        if self.x_reg:
            x_load = "const double *xg = xt + (size_t)tile * TILE_D + lane;\n        " + "\n        ".join(
                "const pm_d3 X%d = pm_d3{xg[%d * 32], xg[%d * 32], xg[%d * 32]};" % (a, 3 * a, 3 * a + 1, 3 * a + 2) for a in range(N))
        else:
            x_load = ""
        pf = 1 if (self.prefetch and self.wf and not self.x_reg) else 0
        pfr = 1 if (pf and self.prefetch == 2) else 0
        import os as _os
        topsync = 1 if (self.sync_every and int(_os.environ.get("PM_SPEC_TOPSYNC", TOP_SYNC))) else 0
        red = 1 if self.fused_sums else 0
        red_params = (", const double *__restrict__ eref, const double *__restrict__ wgt, double d_shift, double *__restrict__ rows"
                      if red else "")
        red_init = "for (int q = 0; q < 5; ++q) ra[q * 32 + lane] = 0.0;" if red else ""
        want_res = ("fres != nullptr || (rows != nullptr && fref_t != nullptr)" if red else "fres != nullptr") if self.wf else "false"
        # fused residual reduction: the five weighted sums of pm_reduce_objective (d = E_ref - E - shift: sum d, sum w d, sum w d^2,
        # sum w, sum w res) are kept per lane in shared memory and become one row of five per warp at the end of the kernel
        red_acc = """if (rows != nullptr) {
            double d = 0.0, ws = 0.0;
            if (live && sconf < n_conf) {
                d = (eref != nullptr ? eref[sconf] : 0.0) - e_total - d_shift;
                ws = wgt != nullptr ? wgt[sconf] : 1.0;
            }
            double *ra_l = ra + lane;
            ra_l[0] += d;
            ra_l[32] = fma(ws, d, ra_l[32]);
            ra_l[64] = fma(ws * d, d, ra_l[64]);
            ra_l[96] += ws;
            ra_l[128] = fma(ws, res, ra_l[128]);
        }""" if red else ""
        red_out = """if (rows != nullptr) {
        __syncwarp();
        double *out = rows + ((size_t)blockIdx.x * n_warps + warp) * 5;
        for (int q = 0; q < 5; ++q) {
            double v = ra[q * 32 + lane];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) out[q] = v;
        }
    }""" if red else ""
        kernel = _KERNEL.format(name=self.name, W=warps, body=body, f_decl=f_decl, res=res, fout=fout, PF=pf, PFR=pfr, TOPSYNC=topsync,
                                RED=red, RED_PARAMS=red_params, RED_INIT=red_init, WANT_RES=want_res, RED_ACC=red_acc, RED_OUT=red_out,
                                f_doubles=0 if (self.f_reg or not self.wf) else 3 * len(self.f_tile_atoms) * 32, x_tile=0 if self.x_reg else 1, x_load=x_load)
        if self.name == "pm_ef_spec_kernel":
            # extra TILE_D buffers per warp of the E+F kernel, read by pm_set_specialized to size the shared memory
            kernel = "__constant__ int spec_ef_info[2] = {%d, %d};\n" % (pfr, red) + kernel
        return kernel if not with_header else _HEADER.format(N=N, ND=self.n_derived) + kernel


_HEADER = r'''// generated by paramol_b200/csrc/specialize.py -- do not edit
#include <cuda_runtime.h>
#include <stdint.h>
#include "pm_device.cuh"

#define SPEC_N {N}
#define SPEC_ND {ND}
#define TILE_D (3 * SPEC_N * 32)
#define PMX(a) pm_d3{{xc_base[(3 * (a)) * 32], xc_base[(3 * (a) + 1) * 32], xc_base[(3 * (a) + 2) * 32]}}
#define PMF_ADD(a, v) do {{ fc_base[(3 * (a)) * 32] += (v).x; fc_base[(3 * (a) + 1) * 32] += (v).y; fc_base[(3 * (a) + 2) * 32] += (v).z; }} while (0)
#define PMF_SUB(a, v) do {{ fc_base[(3 * (a)) * 32] -= (v).x; fc_base[(3 * (a) + 1) * 32] -= (v).y; fc_base[(3 * (a) + 2) * 32] -= (v).z; }} while (0)

// derived table of the parameter vector being evaluated (pm_prepare_params layout), then the 9 N force-metric entries; the
// host copies them here before the launch, so every parameter is a constant-bank operand with an immediate offset
__constant__ double spec_tab[SPEC_ND + 9 * SPEC_N];
'''

_KERNEL = r'''
extern "C" __global__ void __launch_bounds__({W} * 32, 1)
    {name}(const double2 *__restrict__ acos_tab, const double *__restrict__ xt,
           const double *__restrict__ fref_t, long n_conf, long n_tiles, double *__restrict__ emm, double *__restrict__ fres,
           double *__restrict__ fmm_t{RED_PARAMS}) {{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    double2 *s_acos = reinterpret_cast<double2 *>(smem_raw);
    size_t off = ((size_t)(PM_ACOS_N + 1) * sizeof(double2) + 127) & ~(size_t)127;
    const size_t per_warp = 128 + ((size_t)TILE_D * ({x_tile} + {PFR}) + (size_t){f_doubles} + (size_t){RED} * 160) * sizeof(double);
    unsigned char *wbase = smem_raw + off + per_warp * warp;
    uint64_t *bar = reinterpret_cast<uint64_t *>(wbase);
    uint64_t *bar_r = bar + 1;
    double *xs = reinterpret_cast<double *>(wbase + 128);
    double *fs = xs + ({x_tile} ? TILE_D : 0);
    double *rs = fs + {f_doubles};
    double *ra = rs + ({PFR} ? TILE_D : 0);   // fused residual sums: [5][32] running sums of this warp's lanes
    (void)fs;
    (void)rs;
    (void)ra;
    {RED_INIT}
    for (int i = threadIdx.x; i <= PM_ACOS_N; i += blockDim.x) s_acos[i] = acos_tab[i];
    if (lane == 0) {{
        pm_mbar_init(bar, 1);
        pm_mbar_init(bar_r, 1);
        pm_fence_mbar_init();
    }}
    __syncthreads();
    uint32_t parity = 0, parity_r = 0;
    const bool want_res = {WANT_RES};
    const long warp_stride = (long)gridDim.x * n_warps;
    // uniform trip count per CTA (the body may contain CTA barriers): a warp without a tile of its own repeats the last tile
    // of the problem and stores nothing
    const long first_cta = (long)blockIdx.x * n_warps;
    const long n_iter = first_cta < n_tiles ? (n_tiles - first_cta + warp_stride - 1) / warp_stride : 0;
    if ({PF} && n_iter > 0 && lane == 0) {{   // coordinates of the first tile
        long tile0 = first_cta + warp;
        if (tile0 >= n_tiles) tile0 = n_tiles - 1;
        pm_fence_proxy_async();
        pm_mbar_expect_tx(bar, (uint32_t)(TILE_D * sizeof(double)));
        pm_bulk_g2s(xs, xt + (size_t)tile0 * TILE_D, (uint32_t)(TILE_D * sizeof(double)), bar);
    }}
    for (long it = 0; it < n_iter; ++it) {{
        long tile = first_cta + warp + it * warp_stride;
        const bool live = tile < n_tiles;
        if (!live) tile = n_tiles - 1;
        if (lane == 0) {{
            if ({PF}) {{
                // reference forces of this tile into the second buffer while the body runs (the buffer was last read in the
                // previous iteration's epilogue, before its closing __syncwarp)
                if ({PFR} && want_res) {{
                    pm_fence_proxy_async();
                    pm_mbar_expect_tx(bar_r, (uint32_t)(TILE_D * sizeof(double)));
                    pm_bulk_g2s(rs, fref_t + (size_t)tile * TILE_D, (uint32_t)(TILE_D * sizeof(double)), bar_r);
                }} else if (want_res) {{
                    pm_bulk_prefetch_l2(fref_t + (size_t)tile * TILE_D, (uint32_t)(TILE_D * sizeof(double)));
                }}
                if (tile + warp_stride < n_tiles) {{
                    pm_bulk_prefetch_l2(xt + (size_t)(tile + warp_stride) * TILE_D, (uint32_t)(TILE_D * sizeof(double)));
                    if (want_res) pm_bulk_prefetch_l2(fref_t + (size_t)(tile + warp_stride) * TILE_D, (uint32_t)(TILE_D * sizeof(double)));
                }}
            }} else {{
                if ({x_tile}) {{
                    pm_fence_proxy_async();
                    pm_mbar_expect_tx(bar, (uint32_t)(TILE_D * sizeof(double)));
                    pm_bulk_g2s(xs, xt + (size_t)tile * TILE_D, (uint32_t)(TILE_D * sizeof(double)), bar);
                }}
                if (want_res) pm_bulk_prefetch_l2(fref_t + (size_t)tile * TILE_D, (uint32_t)(TILE_D * sizeof(double)));
                if (tile + warp_stride < n_tiles) pm_bulk_prefetch_l2(xt + (size_t)(tile + warp_stride) * TILE_D, (uint32_t)(TILE_D * sizeof(double)));
            }}
        }}
        {x_load}
        {f_decl}
        if ({x_tile}) {{
            pm_mbar_wait(bar, parity);
            parity ^= 1;
        }}
        __syncwarp();
        if ({TOPSYNC}) __syncthreads();   // all warps enter the body together (uniform trip count)
        const double *xc_base = xs + lane;
        double *fc_base = fs + lane;
        (void)fc_base;
        double e_bond = 0.0, e_angle = 0.0, e_tors = 0.0, e_nb = 0.0;
        {{
            {body}
        }}
        if ({PF}) {{
            __syncwarp();   // every lane has read its coordinates: the next tile's may overwrite them during the epilogue
            if (lane == 0 && it + 1 < n_iter) {{
                long next = first_cta + warp + (it + 1) * warp_stride;
                if (next >= n_tiles) next = n_tiles - 1;
                pm_fence_proxy_async();
                pm_mbar_expect_tx(bar, (uint32_t)(TILE_D * sizeof(double)));
                pm_bulk_g2s(xs, xt + (size_t)next * TILE_D, (uint32_t)(TILE_D * sizeof(double)), bar);
            }}
        }}
        const long sconf = tile * 32 + lane;
        const double e_total = ((e_bond + e_angle) + e_tors) + e_nb;
        double res = 0.0;
        if (want_res) {{
            if ({PFR}) {{
                pm_mbar_wait(bar_r, parity_r);
                parity_r ^= 1;
            }}
            const double *fr = {PFR} ? rs + lane : fref_t + (size_t)tile * TILE_D + lane;
            {{
                {res}
            }}
        }}
        {RED_ACC}
        if (live && sconf < n_conf && emm != nullptr) {{
            emm[sconf] = e_total;
            if (fres != nullptr) fres[sconf] = res;
        }}
        if (live && fmm_t != nullptr) {{
            double *fo = fmm_t + (size_t)tile * TILE_D + lane;
            {{
                {fout}
            }}
        }}
        __syncwarp();
    }}
    {RED_OUT}
}}
'''


class GradKernelSource(KernelSource):
    """
    Generated ANALYTIC-GRADIENT kernel (``pm_grad_spec_kernel``): the same blocks as the E+F kernel, one conformation per lane,
    but instead of forces every term produces its contributions to d f / d parameter (see pm_grad_kernels.cu for the algebra:
    a_s dE/dp - (d2E/dq dp)(grad q . lambda)).  The per-lane quantities of 32 terms at a time are summed over the 32
    conformations of the tile with a transposing butterfly (31 shuffle steps for 32 sums) and added to the warp's accumulators
    in shared memory; a second generated kernel adds the warps' partials in a fixed order and scatters them to the gout layout
    of ``pm_grad`` through ``grad_perm``.  Deterministic, no atomics.
    """

    def __init__(self, n_atoms, bonds, angles, torsions, tors_per, pairs, fence_every=2, sync_every=0):
        KernelSource.__init__(self, n_atoms, bonds, angles, torsions, tors_per, pairs, False, "pm_grad_spec_kernel", False, fence_every,
                              sync_every, True)
        self.go_bond = 0
        self.go_angle = 2 * len(self.bonds)
        self.go_tors = self.go_angle + 2 * len(self.angles)
        self.go_pair = self.go_tors + 2 * len(self.torsions)
        self.n_gout = self.go_pair + 3 * len(self.pairs)
        self.order = []   # gout index of every emitted quantity, in emission order

    def L(self, atom):
        return "PML(%d)" % atom

    def quantity(self, gout_index, expr):
        k = len(self.order) % 32
        self.order.append(gout_index)
        self.emit("Q[%d] = %s;" % (k, expr))
        if k == 31:
            self.flush()

    def flush(self, final=False):
        n = len(self.order)
        if final:
            k = n % 32
            if k == 0:
                return
            for i in range(k, 32):
                self.emit("Q[%d] = 0.0;" % i)
                self.order.append(-1)
            n = len(self.order)
        self.emit("{ const double r_ = pmg_reduce32(Q, lane); acc[%d + lane] += r_; }" % (n - 32))

    def gen_stars(self):
        N = self.N
        bond_of = {}
        for t, (i, j) in enumerate(self.bonds):
            bond_of[(i, j)] = t
            bond_of[(j, i)] = t
        angle_at = [[] for _ in range(N)]
        for t, (a, b, c) in enumerate(self.angles):
            angle_at[b].append((t, a, c))
        done_bond = set()
        for c in range(N):
            need = []
            for (i, j) in self.bonds:
                if i == c and j not in need:
                    need.append(j)
                if j == c and i not in need:
                    need.append(i)
            for (_, a, b) in angle_at[c]:
                for v in (a, b):
                    if v not in need:
                        need.append(v)
            own = [n for n in need if (c, n) in bond_of and bond_of[(c, n)] not in done_bond]
            if not own and not angle_at[c]:
                continue
            used = set(own)
            for (_, a, b) in angle_at[c]:
                used.update((a, b))
            need = [n for n in need if n in used]
            self.emit("{  // star %d" % c)
            self.emit("const pm_d3 xc = %s, lc = %s;" % (self.X(c), self.L(c)))
            for n in need:
                self.emit("const pm_d3 d%d = pm_sub(xc, %s); const double q%d = pm_dot(d%d, d%d); const double i%d = pm_rsqrt(q%d);"
                          " const pm_d3 m%d = pm_sub(lc, %s);" % (n, self.X(n), n, n, n, n, n, n, self.L(n)))
            for n in own:
                t = bond_of[(c, n)]
                done_bond.add(t)
                o = self.off_bond + 2 * t
                self.emit("{ const double dl = fma(q%d, i%d, -%s); const double dr = i%d * pm_dot(d%d, m%d);" % (n, n, self.P(o), n, n, n))
                self.quantity(self.go_bond + 2 * t, "%s * (dr - ae * dl)" % self.P(o + 1))
                self.quantity(self.go_bond + 2 * t + 1, "dl * fma(0.5 * ae, dl, -dr)")
                self.emit("}")
            for (t, a, b) in angle_at[c]:
                o = self.off_angle + 2 * t
                self.emit("{ const pm_d3 pc = pm_cross(d%d, d%d); const double pp2 = pm_dot(pc, pc); const double i01 = i%d * i%d;" % (a, b, a, b))
                self.emit("  const double ip = pm_rsqrt(fmax(pp2, 1.0e-12)); const double th = pm_acos_cs(pm_dot(d%d, d%d) * i01, pp2 * ip * i01, s_acos);" % (a, b))
                self.emit("  const double dth = ip * ((i%d * i%d) * pm_dot(pm_cross(d%d, pc), m%d) - (i%d * i%d) * pm_dot(pm_cross(d%d, pc), m%d));"
                          % (a, a, a, a, b, b, b, b))
                self.emit("  const double dl = th - %s;" % self.P(o))
                self.quantity(self.go_angle + 2 * t, "%s * (dth - ae * dl)" % self.P(o + 1))
                self.quantity(self.go_angle + 2 * t + 1, "dl * fma(0.5 * ae, dl, -dth)")
                self.emit("}")
            self.end_block()
        assert len(done_bond) == len(self.bonds)

    def gen_axes(self):
        axes = {}
        for t, (i, j, k, l) in enumerate(self.torsions):
            if (k, j) in axes:
                axes[(k, j)].append((t, l, i))
            else:
                axes.setdefault((j, k), []).append((t, i, l))
        for (j, k), terms in axes.items():
            subs_i = sorted({i for (_, i, _) in terms})
            subs_l = sorted({l for (_, _, l) in terms})
            self.emit("{  // axis %d-%d" % (j, k))
            self.emit("const pm_d3 xj = %s, xk = %s, lj = %s, lk = %s; const pm_d3 d1 = pm_sub(xk, xj); const double bb = pm_dot(d1, d1);"
                      % (self.X(j), self.X(k), self.L(j), self.L(k)))
            self.emit("const double rbb = pm_rsqrt(bb); const double nb = bb * rbb, ibb = rbb * rbb;")
            for i in subs_i:
                self.emit("const pm_d3 a%d = pm_sub(%s, xj); const pm_d3 u%d = pm_cross(a%d, d1); const double s%d = pm_rsqrt(pm_dot(u%d, u%d));"
                          " const double g%d = pm_dot(a%d, d1) * ibb;" % (i, self.X(i), i, i, i, i, i, i, i))
                # w0 = l_i + (ff1 - 1) l_j - ff1 l_k
                self.emit("const pm_d3 li%d = %s; const pm_d3 v%d = pm_d3{li%d.x + (g%d - 1.0) * lj.x - g%d * lk.x, li%d.y + (g%d - 1.0) * lj.y - g%d * lk.y,"
                          " li%d.z + (g%d - 1.0) * lj.z - g%d * lk.z}; const double p%d = pm_dot(u%d, v%d) * (s%d * s%d);"
                          % (i, self.L(i), i, i, i, i, i, i, i, i, i, i, i, i, i, i, i))
            for l in subs_l:
                self.emit("const pm_d3 b%d = pm_sub(xk, %s); const pm_d3 w%d = pm_cross(d1, b%d); const double t%d = pm_rsqrt(pm_dot(w%d, w%d));"
                          " const double h%d = pm_dot(b%d, d1) * ibb;" % (l, self.X(l), l, l, l, l, l, l, l))
                # w3 = l_l - ff2 l_j + (ff2 - 1) l_k
                self.emit("const pm_d3 ll%d = %s; const pm_d3 z%d = pm_d3{ll%d.x - h%d * lj.x + (h%d - 1.0) * lk.x, ll%d.y - h%d * lj.y + (h%d - 1.0) * lk.y,"
                          " ll%d.z - h%d * lj.z + (h%d - 1.0) * lk.z}; const double r%d = pm_dot(w%d, z%d) * (t%d * t%d);"
                          % (l, self.L(l), l, l, l, l, l, l, l, l, l, l, l, l, l, l, l))
            cells = {}
            for (t, i, l) in terms:
                cells.setdefault((i, l), []).append(t)
            for (i, l), ts in cells.items():
                ts = sorted(ts, key=lambda t: self.tors_per[t])
                self.emit("{ const double i12 = s%d * t%d; const double cphi = pm_dot(u%d, w%d) * i12, sphi = nb * pm_dot(a%d, w%d) * i12;" % (i, l, i, l, i, l))
                self.emit("  const double dphi = nb * (p%d - r%d); double cn = 1.0, sn = 0.0;" % (i, l))
                cur = 0
                for t in ts:
                    per = self.tors_per[t]
                    while cur < per:
                        self.emit("  { const double tt = cn * cphi - sn * sphi; sn = fma(sn, cphi, cn * sphi); cn = tt; }")
                        cur += 1
                    o = self.off_tors + 4 * t
                    self.emit("  { const double cpsi = fma(cn, %s, sn * %s), spsi = fma(sn, %s, -(cn * %s));" % (self.P(o + 2), self.P(o + 3), self.P(o + 2), self.P(o + 3)))
                    self.quantity(self.go_tors + 2 * t, "%s * (ae * spsi - %d.0 * cpsi * dphi)" % (self.P(o), per))
                    self.quantity(self.go_tors + 2 * t + 1, "fma(ae, 1.0 + cpsi, %d.0 * spsi * dphi)" % per)
                    self.emit("  }")
                self.emit("}")
            self.end_block()

    def gen_pairs(self):
        by_i = {}
        for q, (i, j, _) in enumerate(self.pairs):
            by_i.setdefault(i, []).append((q, j))
        for i, lst in by_i.items():
            self.emit("{  // pairs of atom %d" % i)
            self.emit("const pm_d3 xi = %s, li = %s;" % (self.X(i), self.L(i)))
            for (q, j) in lst:
                self.emit("{ const pm_d3 d = pm_sub(xi, %s); const double rinv = pm_rsqrt(pm_dot(d, d)); const double u = rinv * rinv; const double u3 = u * u * u;"
                          " const double us = u * pm_dot(d, pm_sub(li, %s));" % (self.X(j), self.L(j)))
                self.quantity(self.go_pair + 3 * q, "(u3 * u3) * fma(12.0, us, ae)")
                self.quantity(self.go_pair + 3 * q + 1, "-u3 * fma(6.0, us, ae)")
                self.quantity(self.go_pair + 3 * q + 2, "rinv * (ae + us)")
                self.emit("}")
            self.end_block()

    def source(self, warps):
        N = self.N
        self.lines = []
        self._blocks = 0
        self.order = []
        self.gen_stars()
        self.gen_axes()
        self.gen_pairs()
        self.flush(final=True)
        body = "\n            ".join(self.lines)
        lam = []
        for a in range(N):
            m = ["spec_tab[SPEC_ND + %d]" % (9 * a + q) for q in range(9)]
            lam.append("{ const double dx = fm[%d * 32] - fr[%d * 32], dy = fm[%d * 32] - fr[%d * 32], dz = fm[%d * 32] - fr[%d * 32];"
                       % (3 * a, 3 * a, 3 * a + 1, 3 * a + 1, 3 * a + 2, 3 * a + 2))
            lam.append("  lc_base[%d * 32] = bf * ((%s + %s) * dx + (%s + %s) * dy + (%s + %s) * dz);" % (3 * a, m[0], m[0], m[1], m[3], m[2], m[6]))
            lam.append("  lc_base[%d * 32] = bf * ((%s + %s) * dx + (%s + %s) * dy + (%s + %s) * dz);" % (3 * a + 1, m[3], m[1], m[4], m[4], m[5], m[7]))
            lam.append("  lc_base[%d * 32] = bf * ((%s + %s) * dx + (%s + %s) * dy + (%s + %s) * dz); }" % (3 * a + 2, m[6], m[2], m[7], m[5], m[8], m[8]))
        ngp = len(self.order)
        perm = ", ".join(str(v) for v in self.order) if self.order else "-1"
        return _GRAD_KERNEL.format(W=warps, body=body, lam="\n                ".join(lam), NGP=max(ngp, 32), perm=perm, NG=self.n_gout)


_GRAD_KERNEL = r'''
#define PML(a) pm_d3{{lc_base[(3 * (a)) * 32], lc_base[(3 * (a) + 1) * 32], lc_base[(3 * (a) + 2) * 32]}}
#define SPEC_NGP {NGP}
__constant__ int grad_perm[SPEC_NGP] = {{{perm}}};
__constant__ int grad_info[2] = {{SPEC_NGP, {NG}}};

// sums of Q[0..31] over the 32 lanes: lane l returns the total of Q[l] (transposing butterfly, 31 shuffle steps)
__device__ __forceinline__ double pmg_reduce32(double (&q)[32], int lane) {{
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {{
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {{
            const double a = q[i], b = q[i + o];
            const double send = up ? a : b, keep = up ? b : a;
            q[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }}
    }}
    return q[0];
}}

extern "C" __global__ void __launch_bounds__({W} * 32, 1)
    pm_grad_spec_kernel(const double2 *__restrict__ acos_tab, const double *__restrict__ xt, const double *__restrict__ fmm_t,
                        const double *__restrict__ fref_t, const double *__restrict__ coef_e, const double *__restrict__ coef_f,
                        long n_conf, long n_tiles, double *__restrict__ partial) {{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    double2 *s_acos = reinterpret_cast<double2 *>(smem_raw);
    size_t off = ((size_t)(PM_ACOS_N + 1) * sizeof(double2) + 127) & ~(size_t)127;
    const size_t per_warp = 128 + (size_t)TILE_D * sizeof(double) * 2 + (size_t)SPEC_NGP * sizeof(double);
    unsigned char *wbase = smem_raw + off + per_warp * warp;
    uint64_t *bar = reinterpret_cast<uint64_t *>(wbase);
    double *xs = reinterpret_cast<double *>(wbase + 128);
    double *ls = xs + TILE_D;
    double *acc = ls + TILE_D;
    for (int i = threadIdx.x; i <= PM_ACOS_N; i += blockDim.x) s_acos[i] = acos_tab[i];
    for (int i = lane; i < SPEC_NGP; i += 32) acc[i] = 0.0;
    if (lane == 0) {{
        pm_mbar_init(bar, 1);
        pm_fence_mbar_init();
    }}
    __syncthreads();
    uint32_t parity = 0;
    const long warp_stride = (long)gridDim.x * n_warps;
    // uniform trip count per CTA (the body may contain CTA barriers): a warp without a tile of its own repeats the last tile with
    // zero coefficients, i.e. adds nothing
    const long first_cta = (long)blockIdx.x * n_warps;
    const long n_iter = first_cta < n_tiles ? (n_tiles - first_cta + warp_stride - 1) / warp_stride : 0;
    for (long it = 0; it < n_iter; ++it) {{
        long tile = first_cta + warp + it * warp_stride;
        const bool live = tile < n_tiles;
        if (!live) tile = n_tiles - 1;
        if (lane == 0) {{
            pm_fence_proxy_async();
            pm_mbar_expect_tx(bar, (uint32_t)(TILE_D * sizeof(double)));
            pm_bulk_g2s(xs, xt + (size_t)tile * TILE_D, (uint32_t)(TILE_D * sizeof(double)), bar);
        }}
        const long sconf = tile * 32 + lane;
        const bool valid = live && sconf < n_conf;
        const double ae = (valid && coef_e != nullptr) ? coef_e[sconf] : 0.0;
        const double bf = (valid && coef_f != nullptr) ? coef_f[sconf] : 0.0;
        double *lc_base = ls + lane;
        if (coef_f != nullptr) {{
            const double *fm = fmm_t + (size_t)tile * TILE_D + lane, *fr = fref_t + (size_t)tile * TILE_D + lane;
            {{
                {lam}
            }}
        }} else {{
            for (int k = 0; k < 3 * SPEC_N; ++k) lc_base[k * 32] = 0.0;
        }}
        pm_mbar_wait(bar, parity);
        parity ^= 1;
        __syncwarp();
        const double *xc_base = xs + lane;
        double Q[32];
        {{
            {body}
        }}
        __syncwarp();
    }}
    double *out = partial + ((size_t)blockIdx.x * n_warps + warp) * SPEC_NGP;
    for (int i = lane; i < SPEC_NGP; i += 32) out[i] = acc[i];
}}

// gout[grad_perm[k]] = sum over the warps' partial rows (fixed order)
extern "C" __global__ void pm_grad_spec_reduce(const double *__restrict__ partial, int n_rows, double *__restrict__ gout) {{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= SPEC_NGP || grad_perm[k] < 0) return;
    double s = 0.0;
    for (int r = 0; r < n_rows; ++r) s += partial[(size_t)r * SPEC_NGP + k];
    gout[grad_perm[k]] = s;
}}
'''


def generate(n_atoms, bonds, angles, torsions, tors_per, pairs, warps, f_in_registers=True, name="pm_ef_spec_kernel",
             x_in_registers=False, fence_every=0, sync_every=0, with_forces=True):
    return KernelSource(n_atoms, bonds, angles, torsions, tors_per, pairs, f_in_registers, name, x_in_registers,
                        fence_every, sync_every, with_forces).source(warps)


# ---------------------------------------------------------------------------------------------------------------------
# compile + cache
# ---------------------------------------------------------------------------------------------------------------------
MAX_SMEM = 227 * 1024
FENCE_EVERY = 2      # measured on B200 (aniline): 2 blocks between compiler fences -> 168 registers, no spills, fastest
SYNC_EVERY = 18      # a CTA barrier every 18 blocks keeps the warps loosely in step, so that fetched instruction lines serve
                     # several of them (measured on B200: aniline 0.630 -> 0.608 ms, 12 and 18 atoms -0.5 ... -2 %; 9 or 14: no gain)
PAIR_SYNC_EVERY = 0      # barriers between the warps of one scheduler only: measured slower (aniline 0.70 ms at 2-3, 0.64 at 6)
ENERGY_SYNC_EVERY = 18   # aniline energy-only: 0 -> 0.293 ms, 12 -> 0.277, 18 -> 0.273
GRAD_SYNC_EVERY = 6   # measured on B200 (aniline gradient): 0 -> 1.04 ms, 4 -> 0.98, 6 -> 0.977, 12 -> 0.995, 18 -> 1.01
MIN_WARPS = 9        # below that the team-mode generic kernel wins (24 atoms: 6 warps -> 0.95x)
MAX_WARPS = 10
MAX_WARPS_ENERGY = 16   # 128 registers, no spills
TOP_SYNC = 0   # CTA barrier at the top of every tile iteration: measured slower (aniline 0.617 -> 0.652 ms: all warps then wait for
               # their coordinate tiles at the same time)
FUSED_SUMS = 0   # residual sums accumulated by the generated kernel itself (rows of five per warp instead of per-conformation energy and
                 # residual + pm_reduce_stage1): measured on B200 with the round-2 body, aniline 2^20: kernel 0.449 -> 0.466 ms with the code
                 # present (register allocation / layout), eval+reduce 0.479 -> 0.497 ms (the reference-energy load at the end of a tile is
                 # exposed with 2 warps per scheduler); 2^17 conformations: 84.0 vs 84.0 us.  Off.


def _red_bytes():
    """Shared memory per warp of the fused residual sums ([5][32] doubles) when that variant is generated."""
    import os
    return 5 * 32 * 8 if int(os.environ.get("PM_SPEC_FUSED", FUSED_SUMS)) else 0


PREFETCH = 0   # 0: coordinates requested at the top of the tile loop (L2-prefetched one tile ahead); 1: next tile's coordinates
               # requested right after the body; 2: also the reference forces by bulk copy into a second tile per warp.
               # Measured on B200 (aniline, 2^20): 0 -> 0.619 ms, 1 -> 0.649 ms, 2 -> 0.636 ms: the waits they remove are
               # already covered by the other warp of the scheduler, the extra warp barrier and 86 KB of shared memory are not


def choose_warps(n_atoms):
    """Warps per CTA of the specialised kernel (coordinate + force tile of 32 conformations per warp), or None when too
    few fit for it to beat the generic kernel."""
    import math
    tile = 3 * n_atoms * 32 * 8
    table = ((257 * 16 + 127) // 128) * 128
    warps = min(MAX_WARPS, int(math.floor((MAX_SMEM - table) / (128 + _red_bytes() + 2 * tile))))
    import os
    return warps if warps >= int(os.environ.get("PM_SPEC_MIN_WARPS", MIN_WARPS)) else None


def _nvcc():
    import os
    import shutil
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def cache_dir():
    import os
    d = os.environ.get("PARAMOL_B200_SPEC_CACHE") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_spec_cache")
    os.makedirs(d, exist_ok=True)
    return d


NVCC_FLAGS = ["-cubin", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xptxas", "-v"]
_NVCC_ID = {}


def _nvcc_identity(nvcc):
    """Release string of the compiler (part of the cache key: another nvcc may generate another image)."""
    import subprocess
    if nvcc not in _NVCC_ID:
        try:
            out = subprocess.run([nvcc, "--version"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True).stdout
            _NVCC_ID[nvcc] = out.strip().splitlines()[-1] if out.strip() else "unknown"
        except OSError:
            _NVCC_ID[nvcc] = "unknown"
    return _NVCC_ID[nvcc]


def compile_image(source, want_report=False):
    """nvcc -cubin for sm_100a, cached by the hash of (source, pm_device.cuh, compiler flags).  The cache directory
    ``paramol_b200/_spec_cache/`` is git-ignored but NOT gpurun-ignored: an image compiled here (e.g. by
    ``__graft_entry__.build()``) travels to the GPU box with the working-tree snapshot and is only loaded there; a molecule first
    seen on a box with nvcc is compiled once on that box.  Safe when several ranks start at once: source, report and image are
    written under per-process names and moved into place with ``os.replace``, so a reader never sees a partial file and
    concurrent writers produce identical bytes (the source is moved into place before it is compiled).  With ``want_report`` also returns the ptxas resource report."""
    import hashlib
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "pm_device.cuh"), "rb") as fh:
        key = hashlib.sha256(source.encode() + fh.read() + " ".join(NVCC_FLAGS).encode()).hexdigest()[:24]
    path = os.path.join(cache_dir(), key + ".cubin")
    if not os.path.exists(path) or not os.path.exists(path + ".txt"):
        nvcc = _nvcc()
        if nvcc is None:
            raise RuntimeError("nvcc not found: cannot compile the topology-specialised kernel")
        tag = ".tmp%d" % os.getpid()
        final_src = os.path.join(cache_dir(), key + ".cu")
        src = final_src + tag
        with open(src, "w") as fh:
            fh.write(source)
        # the source goes into place first (atomically; concurrent ranks write identical bytes) and is compiled under its final
        # name, so that the -lineinfo records of the image point at a file that exists (ncu source view)
        os.replace(src, final_src)
        tmp = path + tag
        try:
            res = subprocess.run([nvcc] + NVCC_FLAGS + ["-I", here, final_src, "-o", tmp], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                 universal_newlines=True)
            if res.returncode != 0:
                raise RuntimeError("nvcc failed on the specialised kernel:\n" + res.stdout[-2000:])
            with open(tmp + ".txt", "w") as fh:
                fh.write("# %s\n# %s\n" % (_nvcc_identity(nvcc), " ".join(NVCC_FLAGS)) + res.stdout)
            os.replace(tmp + ".txt", path + ".txt")
            os.replace(tmp, path)
        finally:
            for leftover in (src, tmp, tmp + ".txt"):
                if os.path.exists(leftover):
                    os.remove(leftover)
    with open(path, "rb") as fh:
        image = fh.read()
    if want_report:
        with open(path + ".txt") as fh:
            return image, fh.read()
    return image


def _spills(report, kernel):
    """Bytes of spill stores ptxas reports for ``kernel`` (0 when it fits its register budget)."""
    lines = report.splitlines()
    for i, line in enumerate(lines):
        if "Compiling entry function '%s'" % kernel in line:
            for nxt in lines[i + 1:i + 4]:
                if "spill stores" in nxt:
                    return int(nxt.split("bytes stack frame,")[1].split("bytes spill stores")[0])
    return -1


def choose_warps_grad(n_atoms, n_gout):
    """Warps per CTA of the generated gradient kernel (coordinate tile + lambda tile + the warp's accumulators), or 0."""
    import math
    tile = 3 * n_atoms * 32 * 8
    table = ((257 * 16 + 127) // 128) * 128
    ngp = max(32, (n_gout + 31) // 32 * 32)
    warps = min(MAX_WARPS, int(math.floor((MAX_SMEM - table) / (128 + 2 * tile + 8 * ngp))))
    return warps if warps >= 6 else 0


def choose_warps_hybrid(n_atoms, n_tile_atoms):
    """Warps per CTA when n_tile_atoms keep their forces in a packed shared tile and the rest in registers (255 registers: 8)."""
    import math
    table = ((257 * 16 + 127) // 128) * 128
    per_warp = 128 + _red_bytes() + 3 * (n_atoms + n_tile_atoms) * 32 * 8
    warps = min(F_REG_WARPS, int(math.floor((MAX_SMEM - table) / per_warp)))
    return warps if warps >= F_REG_MIN_WARPS else None


def choose_warps_energy(n_atoms):
    """Warps per CTA of the energy-only specialised kernel (coordinate tile only)."""
    import math
    tile = 3 * n_atoms * 32 * 8
    table = ((257 * 16 + 127) // 128) * 128
    warps = min(MAX_WARPS_ENERGY, int(math.floor((MAX_SMEM - table) / (128 + _red_bytes() + tile))))
    return warps if warps >= MIN_WARPS else None


F_REG_WARPS = 8       # forces in registers: 255 registers per thread
F_REG_MIN_WARPS = 6
F_REG_MAX_ATOMS = 26   # measured on B200 vs the team-mode interpreter with the round-2 body (scripts/spec_size_sweep.py, 2^18 conformations,
                       # M conformations/s, generated : interpreter): 20 atoms 727 : 529, 22 atoms 652 : 510, caffeine (24) 426 : 374,
                       # 26 atoms 390 : 360, 28 atoms 296 : 299, 32 atoms 186 : 314 (the force-tile shape with 4-7 warps never wins there)
F_REG_MAX_SPILL = 512  # bytes of spill stores tolerated (aniline at fence 6: 204 bytes, still the fastest)


def f_reg_fence(n_atoms):
    """Blocks between scheduling fences for the forces-in-registers shape: the more atoms, the fewer registers are left for
    interleaving (measured on B200, round-1 body: aniline 4 -> 0.646 ms, 6 -> 0.632 ms, 8 -> 0.659 ms, 10 -> 0.714 ms with spills; final
    round-2 body, which needs fewer registers (profiles/r3_spec_sweep.txt, barrier every 18 blocks): 4 -> 0.442, 6 -> 0.444, 8 -> 0.437,
    9 -> 0.453 ms; without the barrier 0.58-0.62 ms: the loose lock step of the warps is what keeps the body in the instruction cache;
    20 atoms 2 -> 727,
    1 -> 717 M conformations/s; 22 atoms 2 -> 578, 1 -> 652; caffeine 2 -> 409, 1 -> 426; 26 atoms 2 -> 282, 1 -> 390)."""
    return 8 if n_atoms <= 14 else (2 if n_atoms <= 20 else 1)


def choose_warps_f_reg(n_atoms):
    """Warps per CTA of the forces-in-registers shape (coordinate tile only in shared memory; 255 registers cap it at 8)."""
    import math
    import os
    if n_atoms > int(os.environ.get("PM_SPEC_FREG_MAX", F_REG_MAX_ATOMS)):
        return None
    tile = 3 * n_atoms * 32 * 8
    table = ((257 * 16 + 127) // 128) * 128
    warps = min(F_REG_WARPS, int(math.floor((MAX_SMEM - table) / (128 + _red_bytes() + tile))))
    return warps if warps >= F_REG_MIN_WARPS else None


def eligible(n_atoms):
    return choose_warps_f_reg(n_atoms) is not None or choose_warps(n_atoms) is not None


def image_for(n_atoms, bonds, angles, torsions, tors_per, pairs):
    """
    One cubin with ``pm_ef_spec_kernel`` (energies, forces, residual) and, when enough warps fit, ``pm_e_spec_kernel``
    (energies only) and the gradient kernels sharing ``spec_tab``.  Returns (image, warps, atoms in the force tile, warps_energy,
    warps_grad) or None when neither shape of
    the E+F kernel suits the molecule.

    Two shapes are tried: forces in REGISTERS (3 N doubles per lane, 8 warps, no force tile: fewer instructions and no
    shared-memory read-modify-write chains; aniline 0.632 ms) is kept when ptxas reports (almost) no spills -- up to ~24 atoms;
    otherwise forces live in the shared force tile (10 warps, 168 registers; aniline 0.704 ms) if at least 9 warps fit.
    """
    import os
    sync = int(os.environ.get("PM_SPEC_SYNC", SYNC_EVERY))
    we = choose_warps_energy(n_atoms) or 0

    grad = GradKernelSource(n_atoms, bonds, angles, torsions, tors_per, pairs, fence_every=FENCE_EVERY,
                            sync_every=int(os.environ.get("PM_SPEC_GRAD_SYNC", GRAD_SYNC_EVERY))) \
        if os.environ.get("PM_SPEC_GRAD", "1") != "0" else None
    wg = choose_warps_grad(n_atoms, grad.n_gout) if grad is not None else 0
    if wg and os.environ.get("PM_SPEC_GRAD_WARPS"):
        wg = min(wg, int(os.environ["PM_SPEC_GRAD_WARPS"]))
    if grad is not None and os.environ.get("PM_SPEC_GRAD_FENCE"):
        grad.fence_every = int(os.environ["PM_SPEC_GRAD_FENCE"])

    def build(f_reg, w, fence):
        ks = KernelSource(n_atoms, bonds, angles, torsions, tors_per, pairs, f_reg, "pm_ef_spec_kernel", False, fence, sync)
        tile = 3 * n_atoms * 32 * 8
        table = ((257 * 16 + 127) // 128) * 128
        ks.prefetch = int(os.environ.get("PM_SPEC_PREFETCH", PREFETCH))
        if ks.prefetch == 2 and table + w * (128 + _red_bytes() + 2 * tile + 3 * len(ks.f_tile_atoms) * 32 * 8) > MAX_SMEM:
            ks.prefetch = 1
        src = ks.source(w)
        if we:
            src += KernelSource(n_atoms, bonds, angles, torsions, tors_per, pairs, False, "pm_e_spec_kernel", False, FENCE_EVERY,
                                int(os.environ.get("PM_SPEC_E_SYNC", ENERGY_SYNC_EVERY)), with_forces=False).source(we, with_header=False)
        if wg:
            src += grad.source(wg)
        return compile_image(src, want_report=True)

    mode = os.environ.get("PM_SPEC_FREG", "auto")
    w_tile = choose_warps(n_atoms)

    def busiest(k):
        count = np.zeros(n_atoms)
        for arr in (bonds, angles, torsions):
            for a in np.asarray(arr).ravel():
                count[int(a)] += 1
        return [int(a) for a in np.argsort(-count, kind="stable")[:k]]

    if mode.startswith("top"):   # experiment: only the k busiest atoms in registers, the others in a packed force tile
        keep = busiest(int(mode[3:]))
        w = int(os.environ.get("PM_SPEC_WARPS", choose_warps_hybrid(n_atoms, n_atoms - len(keep)) or 0))
        if w < 1:
            return None
        image, _ = build(keep, w, int(os.environ.get("PM_SPEC_FENCE", FENCE_EVERY)))
        return image, w, n_atoms - len(keep), we, wg
    w_reg = choose_warps_f_reg(n_atoms)
    if mode != "0" and w_reg:
        w = int(os.environ.get("PM_SPEC_WARPS", w_reg))
        image, report = build(True, w, int(os.environ.get("PM_SPEC_FENCE", f_reg_fence(n_atoms))))
        if 0 <= _spills(report, "pm_ef_spec_kernel") <= F_REG_MAX_SPILL or mode == "1":
            return image, w, 0, we, wg
    if w_tile:
        w = int(os.environ.get("PM_SPEC_WARPS", w_tile))
        image, _ = build(False, w, int(os.environ.get("PM_SPEC_FENCE", FENCE_EVERY)))
        return image, w, n_atoms, we, wg
    return None
