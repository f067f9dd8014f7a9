#!/usr/bin/env python3
"""Generates qcs_regops.inc: the register-op interpreter of the tile engine as ONE inline-PTX block per
amplitude type (complex64: 2^4 amplitudes per thread, complex128: 2^3).

Why PTX and not C++: a fused pass is instruction-issue-bound.  Written as a C++ `switch` over ~60 arms the
compiler (a) lowers the dispatch to a 6-level compare/branch tree (12 instructions and ~6 dependent branch
latencies per op) and (b) decodes every field of the op record before the switch.  ncu on the C++ version:
BRA + ISETP = 23 % of the instructions, FP32 arithmetic only 29 %.  Here the dispatch is one `brx.idx`
through a jump table, each arm decodes only what it needs, and every arm updates the amplitude registers
in place.

Op record (16 bytes, shared memory): w0 = arm | reg_care << 16 | reg_val << 24, w1 = base_care | base_val << 16,
w2 = byte offset of the coefficients in the pool, w3 = tile-level predicate slot (63 = none, always true).
The arm table (ARMS below) is shared with the host planner through the generated constants.

Run:  python3 gen_regops.py > qcs_regops.inc     (the Makefile does this)
"""
import sys


class Gen:
    def __init__(self, ty, rb, with_gen2=True):
        self.with_gen2 = with_gen2        # False: the lean interpreter (GEN2 ids jump to the loop head)
        self.ty = ty                      # 'f32' | 'f64'
        self.rb = rb
        self.nx = 1 << rb
        self.lines = []
        self.labels = []                  # jump table, index = arm id
        self.uid = 0
        self.namp = 2 * self.nx           # real registers
        self.PH = (self.namp, self.namp + 1)
        self.OPS, self.POOL, self.BASE, self.OUTER = (self.namp + 2, self.namp + 3, self.namp + 4, self.namp + 5)
        self.esz = 4 if ty == 'f32' else 8

    # operand names
    def re(self, x): return '%%%d' % (2 * x)
    def im(self, x): return '%%%d' % (2 * x + 1)

    def emit(self, s): self.lines.append(s)

    def tmp(self):
        self.uid += 1
        return self.uid

    def label(self, name):
        self.emit('%s:' % name)

    # ---- primitives (in place) ----
    def map2(self, u, v, p, q, r, s):
        t = self.tmp()
        self.emit('mul.%s m1_%d, %s, %s;' % (self.ty, t, q, v))
        self.emit('mul.%s m2_%d, %s, %s;' % (self.ty, t, r, u))
        self.emit('fma.rn.%s %s, %s, %s, m1_%d;' % (self.ty, u, p, u, t))
        self.emit('fma.rn.%s %s, %s, %s, m2_%d;' % (self.ty, v, s, v, t))
        self.temps.add(('m1_%d' % t)); self.temps.add(('m2_%d' % t))

    def map2u(self, u, v, q, r, d):
        """u' = u + q v ; v' = v + r u = d v + r u' with d = 1 - r q (no copy of the old u is needed)"""
        self.emit('fma.rn.%s %s, %s, %s, %s;' % (self.ty, u, q, v, u))
        self.emit('mul.%s %s, %s, %s;' % (self.ty, v, v, d))
        self.emit('fma.rn.%s %s, %s, %s, %s;' % (self.ty, v, r, u, v))

    def addsub(self, u, v):
        m2 = '0fC0000000' if self.ty == 'f32' else '0dC000000000000000'
        self.emit('add.%s %s, %s, %s;' % (self.ty, u, u, v))
        self.emit('fma.rn.%s %s, %s, %s, %s;' % (self.ty, v, v, m2, u))

    def gen2(self, x0, x1):
        """(t0, t1) -> (m00 t0 + m01 t1, m10 t0 + m11 t1); coefficients in c0..c7 (m00 m01 m10 m11 as re, im),
        negated imaginary parts in n1 n3 n5 n7."""
        t = self.tmp()
        ty = self.ty
        a0x, a0y, a1x, a1y = self.re(x0), self.im(x0), self.re(x1), self.im(x1)
        p = ['g%d_%d' % (i, t) for i in range(4)]     # p0x p0y p1x p1y
        for n in p:
            self.temps.add(n)
        e = self.emit
        e('mul.%s %s, c0, %s;' % (ty, p[0], a0x)); e('fma.rn.%s %s, n1, %s, %s;' % (ty, p[0], a0y, p[0]))
        e('mul.%s %s, c0, %s;' % (ty, p[1], a0y)); e('fma.rn.%s %s, c1, %s, %s;' % (ty, p[1], a0x, p[1]))
        e('mul.%s %s, c4, %s;' % (ty, p[2], a0x)); e('fma.rn.%s %s, n5, %s, %s;' % (ty, p[2], a0y, p[2]))
        e('mul.%s %s, c4, %s;' % (ty, p[3], a0y)); e('fma.rn.%s %s, c5, %s, %s;' % (ty, p[3], a0x, p[3]))
        e('fma.rn.%s %s, c2, %s, %s;' % (ty, p[0], a1x, p[0])); e('fma.rn.%s %s, n3, %s, %s;' % (ty, a0x, a1y, p[0]))
        e('fma.rn.%s %s, c2, %s, %s;' % (ty, p[1], a1y, p[1])); e('fma.rn.%s %s, c3, %s, %s;' % (ty, a0y, a1x, p[1]))
        e('fma.rn.%s %s, c7, %s, %s;' % (ty, p[3], a1x, p[3])); e('fma.rn.%s %s, n7, %s, %s;' % (ty, p[2], a1y, p[2]))
        e('fma.rn.%s %s, c6, %s, %s;' % (ty, a1x, a1x, p[2])); e('fma.rn.%s %s, c6, %s, %s;' % (ty, a1y, a1y, p[3]))

    # ---- packed FP32 (complex64 only): the real and imaginary part of one amplitude share a 64-bit register
    # pair, so arms that apply the same real coefficients to both halves issue half the FP instructions
    # (FADD2 / FMUL2 / FFMA2).  The pack / unpack moves vanish: the amplitudes already live in aligned pairs
    # (they arrive through 64-bit shared-memory loads).
    def packed(self):
        return self.ty == 'f32'

    def pk(self, x):
        """packed register name of amplitude x (declared per arm by pack())"""
        return 'P%d' % x

    def pack(self, xs):
        for x in xs:
            self.emit('mov.b64 %s, {%s, %s};' % (self.pk(x), self.re(x), self.im(x)))

    def unpack(self, xs):
        for x in xs:
            self.emit('mov.b64 {%s, %s}, %s;' % (self.re(x), self.im(x), self.pk(x)))

    def bcast(self, dst, src):
        self.emit('mov.b64 %s, {%s, %s};' % (dst, src, src))

    # ---- coefficient loads ----
    def ld4(self, regs=('c0', 'c1', 'c2', 'c3'), off=0):
        if self.ty == 'f32':
            self.emit('ld.shared.v4.f32 {%s, %s, %s, %s}, [cf+%d];' % (regs + (off,)))
        else:
            self.emit('ld.shared.v2.f64 {%s, %s}, [cf+%d];' % (regs[0], regs[1], off))
            self.emit('ld.shared.v2.f64 {%s, %s}, [cf+%d];' % (regs[2], regs[3], off + 16))

    def ld2(self, regs=('c0', 'c1')):
        self.emit('ld.shared.v2.%s {%s, %s}, [cf];' % (self.ty, regs[0], regs[1]))

    def ld1(self, reg='c0'):
        self.emit('ld.shared.%s %s, [cf];' % (self.ty, reg))

    def cfaddr(self):
        self.emit('add.u32 cf, w2, %%%d;' % self.POOL)

    def compute_ok(self):
        e = self.emit
        e('ld.shared.v4.u32 {k0, k1, k2, k3}, [opa+-32];')     # this op's record (w* already hold the next one)
        e('and.b32 t0, k1, 0xffff;')
        e('shr.u32 t1, k1, 16;')
        e('and.b32 t2, %%%d, t0;' % self.BASE)
        e('setp.eq.u32 ok, t2, t1;')
        e('shr.u64 tt, %%%d, k3;' % self.OUTER)
        e('cvt.u32.u64 t3, tt;')
        e('and.b32 t3, t3, 1;')
        e('setp.ne.u32 pt, t3, 0;')
        e('and.pred ok, ok, pt;')

    def one(self): return '0f3F800000' if self.ty == 'f32' else '0d3FF0000000000000'
    def zero(self): return '0f00000000' if self.ty == 'f32' else '0d0000000000000000'

    def sel_identity4(self):
        e = self.emit
        e('selp.%s c0, c0, %s, ok;' % (self.ty, self.one()))
        e('selp.%s c1, c1, %s, ok;' % (self.ty, self.zero()))
        e('selp.%s c2, c2, %s, ok;' % (self.ty, self.zero()))
        e('selp.%s c3, c3, %s, ok;' % (self.ty, self.one()))

    def rcrv(self):
        self.emit('bfe.u32 rc, k0, 16, 8;')
        self.emit('shr.u32 rv, k0, 24;')

    def done_arm(self):
        self.emit('bra.uni L_NEXT;')

    def pairs(self, j, extra=0):
        return [(x0, x0 | (1 << j)) for x0 in range(self.nx) if not (x0 >> j) & 1 and (x0 & extra) == extra]

    # ---- arms ----
    def arm(self, name):
        self.labels.append(name)
        self.label(name)

    def build(self):
        self.temps = set()
        self.temps64 = set()
        rb, nx, ty = self.rb, self.nx, self.ty
        body_start = len(self.lines)
        J = range(4)

        def absent(name):
            # arm ids are shared by both amplitude types; positions >= rb do not exist for this one
            self.labels.append('L_NEXT')

        # HAD
        for j in J:
            if j >= rb: absent('HAD'); continue
            self.arm('A_HAD%d' % j)
            if self.packed():
                self.pack(range(nx))
                for x0, x1 in self.pairs(j):
                    self.emit('add.rn.f32x2 %s, %s, %s;' % (self.pk(x0), self.pk(x0), self.pk(x1)))
                    self.emit('fma.rn.f32x2 %s, %s, QM2, %s;' % (self.pk(x1), self.pk(x1), self.pk(x0)))
                self.unpack(range(nx))
            else:
                for x0, x1 in self.pairs(j):
                    self.addsub(self.re(x0), self.re(x1)); self.addsub(self.im(x0), self.im(x1))
            self.done_arm()
        # REAL (plain, then predicated entry)
        def real_family(pred):
            for j in J:
                if j >= rb: absent('REAL'); continue
                self.arm('A_REAL%s%d' % ('P' if pred else '', j))
                if pred:
                    self.compute_ok(); self.sel_identity4()
                    self.emit('bra.uni B_REAL%d;' % j)
                else:
                    self.label('B_REAL%d' % j)
                    if self.packed():
                        for k in range(4):
                            self.bcast('Q%d' % k, 'c%d' % k)
                        self.pack(range(nx))
                        for x0, x1 in self.pairs(j):
                            t = self.tmp()
                            self.temps64.add('pm1_%d' % t); self.temps64.add('pm2_%d' % t)
                            self.emit('mul.rn.f32x2 pm1_%d, Q1, %s;' % (t, self.pk(x1)))
                            self.emit('mul.rn.f32x2 pm2_%d, Q2, %s;' % (t, self.pk(x0)))
                            self.emit('fma.rn.f32x2 %s, Q0, %s, pm1_%d;' % (self.pk(x0), self.pk(x0), t))
                            self.emit('fma.rn.f32x2 %s, Q3, %s, pm2_%d;' % (self.pk(x1), self.pk(x1), t))
                        self.unpack(range(nx))
                    else:
                        for x0, x1 in self.pairs(j):
                            self.map2(self.re(x0), self.re(x1), 'c0', 'c1', 'c2', 'c3')
                            self.map2(self.im(x0), self.im(x1), 'c0', 'c1', 'c2', 'c3')
                    self.done_arm()
        real_family(False)
        real_family(True)
        # REALU
        for j in J:
            if j >= rb: absent('REALU'); continue
            self.arm('A_REALU%d' % j)
            if self.packed():
                for k in range(3):
                    self.bcast('Q%d' % k, 'c%d' % k)
                self.pack(range(nx))
                for x0, x1 in self.pairs(j):
                    self.emit('fma.rn.f32x2 %s, Q0, %s, %s;' % (self.pk(x0), self.pk(x1), self.pk(x0)))
                    self.emit('mul.rn.f32x2 %s, %s, Q2;' % (self.pk(x1), self.pk(x1)))
                    self.emit('fma.rn.f32x2 %s, Q1, %s, %s;' % (self.pk(x1), self.pk(x0), self.pk(x1)))
                self.unpack(range(nx))
            else:
                for x0, x1 in self.pairs(j):
                    self.map2u(self.re(x0), self.re(x1), 'c0', 'c1', 'c2'); self.map2u(self.im(x0), self.im(x1), 'c0', 'c1', 'c2')
            self.done_arm()
        # IMAG
        def imag_family(pred):
            for j in J:
                if j >= rb: absent('IMAG'); continue
                self.arm('A_IMAG%s%d' % ('P' if pred else '', j))
                if pred:
                    self.compute_ok(); self.sel_identity4()
                    self.emit('bra.uni B_IMAG%d;' % j)
                else:
                    self.label('B_IMAG%d' % j)
                    self.emit('neg.%s n1, c1;' % ty); self.emit('neg.%s n3, c2;' % ty)
                    for x0, x1 in self.pairs(j):
                        self.map2(self.re(x0), self.im(x1), 'c0', 'c1', 'c2', 'c3')
                        self.map2(self.im(x0), self.re(x1), 'c0', 'n1', 'n3', 'c3')
                    self.done_arm()
        imag_family(False)
        imag_family(True)
        # IMAGU
        for j in J:
            if j >= rb: absent('IMAGU'); continue
            self.arm('A_IMAGU%d' % j)
            self.emit('neg.%s n1, c0;' % ty); self.emit('neg.%s n3, c1;' % ty)
            for x0, x1 in self.pairs(j):
                self.map2u(self.re(x0), self.im(x1), 'c0', 'c1', 'c2'); self.map2u(self.im(x0), self.re(x1), 'n1', 'n3', 'c2')
            self.done_arm()
        # GEN (plain / predicated)
        def gen_coefs(pred):
            self.ld4(('c4', 'c5', 'c6', 'c7'), 4 * self.esz)
            if pred:
                e = self.emit
                for r, idv in (('c0', self.one()), ('c1', self.zero()), ('c2', self.zero()), ('c3', self.zero()),
                               ('c4', self.zero()), ('c5', self.zero()), ('c6', self.one()), ('c7', self.zero())):
                    e('selp.%s %s, %s, %s, ok;' % (ty, r, r, idv))
        def gen_negs():
            for i in (1, 3, 5, 7):
                self.emit('neg.%s n%d, c%d;' % (ty, i, i))
        def gen_family(pred):
            for j in J:
                if j >= rb: absent('GEN'); continue
                self.arm('A_GEN%s%d' % ('P' if pred else '', j))
                if pred:
                    self.compute_ok()
                gen_coefs(pred)
                if pred:
                    self.emit('bra.uni B_GEN%d;' % j)
                else:
                    self.label('B_GEN%d' % j)
                    gen_negs()
                    for x0, x1 in self.pairs(j):
                        self.gen2(x0, x1)
                    self.done_arm()
        gen_family(False)
        gen_family(True)
        # GENM: register-bit control pattern (rc, rv), always predicated
        for j in J:
            if j >= rb: absent('GENM'); continue
            self.arm('A_GENM%d' % j)
            self.compute_ok(); gen_coefs(True); gen_negs(); self.rcrv()
            for x0, x1 in self.pairs(j):
                t = self.tmp()
                self.emit('and.b32 t0, rc, %d;' % x0)
                self.emit('setp.ne.u32 pt, t0, rv;')
                self.emit('@pt bra.uni S_GENM_%d;' % t)
                self.gen2(x0, x1)
                self.label('S_GENM_%d' % t)
            self.done_arm()
        # CX: X on bit j where register bit cb is 1
        for j in J:
            for cb in J:
                if cb == j: continue
                if j >= rb or cb >= rb: absent('CX'); continue
                self.arm('A_CX%d_%d' % (j, cb))
                for x0, x1 in self.pairs(j, 1 << cb):
                    for a, b in ((self.re(x0), self.re(x1)), (self.im(x0), self.im(x1))):
                        self.emit('mov.%s sw, %s;' % (ty, a)); self.emit('mov.%s %s, %s;' % (ty, a, b)); self.emit('mov.%s %s, sw;' % (ty, b))
                self.done_arm()
        # phases on the amplitudes whose register bits in `mask` are all 1
        def phase_family(tag, masks, pred):
            for k, mask in enumerate(masks):
                if mask is None: absent(tag); continue
                self.arm('A_%s%s%d' % (tag, 'P' if pred else '', k))
                if pred:
                    self.compute_ok(); self.sel_identity4()
                    self.emit('bra.uni B_%s%d;' % (tag, k))
                else:
                    self.label('B_%s%d' % (tag, k))
                    for x in range(nx):
                        if (x & mask) == mask:
                            self.map2(self.re(x), self.im(x), 'c0', 'c1', 'c2', 'c3')
                    self.done_arm()
        def sign_family(tag, masks, pred):
            for k, mask in enumerate(masks):
                if mask is None: absent(tag); continue
                self.arm('A_%s%s%d' % (tag, 'P' if pred else '', k))
                if pred:
                    self.compute_ok()
                    self.emit('selp.%s c0, c0, %s, ok;' % (ty, self.one()))
                    self.emit('bra.uni B_%s%d;' % (tag, k))
                else:
                    self.label('B_%s%d' % (tag, k))
                    sel = [x for x in range(nx) if (x & mask) == mask]
                    if self.packed():
                        self.bcast('Q0', 'c0')
                        self.pack(sel)
                        for x in sel:
                            self.emit('mul.rn.f32x2 %s, %s, Q0;' % (self.pk(x), self.pk(x)))
                        self.unpack(sel)
                    else:
                        for x in sel:
                            self.emit('mul.%s %s, %s, c0;' % (ty, self.re(x), self.re(x)))
                            self.emit('mul.%s %s, %s, c0;' % (ty, self.im(x), self.im(x)))
                    self.done_arm()
        m1 = [(1 << j) if j < rb else None for j in J]
        prs = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
        m2 = [((1 << a) | (1 << b)) if b < rb else None for a, b in prs]
        phase_family('PH1', m1, False); phase_family('PH1', m1, True)
        sign_family('SGN1', m1, False); sign_family('SGN1', m1, True)
        phase_family('PH2', m2, False); phase_family('PH2', m2, True)
        sign_family('SGN2', m2, False); sign_family('SGN2', m2, True)
        # PHT: thread-uniform phase (always predicated: a uniform unpredicated phase is the pass-level factor)
        self.arm('A_PHT')
        self.compute_ok(); self.sel_identity4()
        self.map2('%%%d' % self.PH[0], '%%%d' % self.PH[1], 'c0', 'c1', 'c2', 'c3')
        self.done_arm()
        # PHM: phase on any (care, value) pattern over the register bits
        self.arm('A_PHM')
        self.compute_ok(); self.sel_identity4(); self.rcrv()
        for x in range(nx):
            t = self.tmp()
            self.emit('and.b32 t0, rc, %d;' % x)
            self.emit('setp.ne.u32 pt, t0, rv;')
            self.emit('@pt bra.uni S_PHM_%d;' % t)
            self.map2(self.re(x), self.im(x), 'c0', 'c1', 'c2', 'c3')
            self.label('S_PHM_%d' % t)
        self.done_arm()
        # END of the sweep's op list
        self.arm('A_END')
        self.emit('bra.uni L_DONE;')
        # GEN2: any complex 4x4 on two register bits (pair p = (j1 < j2), matrix index = bit j2 : bit j1), in place:
        # the off-diagonal terms of all four outputs go to temporaries first, then each amplitude is updated by
        # its own diagonal entry (so no copy back).  32 coefficients e0..e31 (row-major, re / im).
        def gen2_groups(j1, j2):
            out = []
            for x0 in range(nx):
                if (x0 >> j1) & 1 or (x0 >> j2) & 1:
                    continue
                out.append([x0 | ((i & 1) << j1) | (((i >> 1) & 1) << j2) for i in range(4)])
            return out

        def gen2_row(y, xs, T, coef):
            """T[y] = sum_{x != y} M[y][x] a_x + i Im(M[y][y]) a_y ; coef(y, x) -> (re, im) register names"""
            tx, ty_ = T[y]
            first = True
            for x in range(4):
                if x == y:
                    continue
                mr, mi = coef(y, x)
                ax, ay = self.re(xs[x]), self.im(xs[x])
                if first:
                    self.emit('mul.%s %s, %s, %s;' % (ty, tx, mr, ax))
                    self.emit('mul.%s %s, %s, %s;' % (ty, ty_, mr, ay))
                    first = False
                else:
                    self.emit('fma.rn.%s %s, %s, %s, %s;' % (ty, tx, mr, ax, tx))
                    self.emit('fma.rn.%s %s, %s, %s, %s;' % (ty, ty_, mr, ay, ty_))
                self.emit('neg.%s ngt, %s;' % (ty, mi))      # single-use negation: folds into the FMA operand
                self.emit('fma.rn.%s %s, ngt, %s, %s;' % (ty, tx, ay, tx))
                self.emit('fma.rn.%s %s, %s, %s, %s;' % (ty, ty_, mi, ax, ty_))
            di = coef(y, y)[1]
            self.emit('neg.%s ngt, %s;' % (ty, di))
            self.emit('fma.rn.%s %s, ngt, %s, %s;' % (ty, tx, self.im(xs[y]), tx))
            self.emit('fma.rn.%s %s, %s, %s, %s;' % (ty, ty_, di, self.re(xs[y]), ty_))

        def gen2_temps(ngroups):
            t = self.tmp()
            T = [[('h%d_%dx_%d' % (g, y, t), 'h%d_%dy_%d' % (g, y, t)) for y in range(4)] for g in range(ngroups)]
            for g in T:
                for a, b in g:
                    self.temps.add(a); self.temps.add(b)
            return T

        def gen2_family(pred):
            # complex64: all 32 coefficients in registers, the predicated entry substitutes the identity and joins
            # the plain body.  complex128 would spill that way: it loads one matrix row at a time and both entries
            # run the same body with `ok` (always true for the plain entry) selecting the row or the identity row.
            rowwise = ty == 'f64'
            for k, (j1, j2) in enumerate(prs):
                if j2 >= rb or not self.with_gen2: absent('GEN2'); continue
                self.arm('A_GEN2%s%d' % ('P' if pred else '', k))
                groups = gen2_groups(j1, j2)
                if rowwise:
                    if pred:
                        self.compute_ok()
                        self.emit('bra.uni B_GEN2_%d;' % k)
                        continue
                    self.emit('setp.eq.u32 ok, 0, 0;')
                    self.label('B_GEN2_%d' % k)
                    T = gen2_temps(len(groups))
                    for y in range(4):
                        self.ld4(('e0', 'e1', 'e2', 'e3'), 8 * y * self.esz)
                        self.ld4(('e4', 'e5', 'e6', 'e7'), (8 * y + 4) * self.esz)
                        for i in range(8):
                            self.emit('selp.%s e%d, e%d, %s, ok;' % (ty, i, i, self.one() if i == 2 * y else self.zero()))
                        for g, xs in enumerate(groups):
                            gen2_row(y, xs, T[g], lambda yy, x: ('e%d' % (2 * x), 'e%d' % (2 * x + 1)))
                        self.emit('mov.%s e%d, e%d;' % (ty, 8 + y, 2 * y))          # keep Re M[y][y] for the final step
                    for g, xs in enumerate(groups):
                        for y in range(4):
                            self.emit('fma.rn.%s %s, e%d, %s, %s;' % (ty, self.re(xs[y]), 8 + y, self.re(xs[y]), T[g][y][0]))
                            self.emit('fma.rn.%s %s, e%d, %s, %s;' % (ty, self.im(xs[y]), 8 + y, self.im(xs[y]), T[g][y][1]))
                    self.done_arm()
                    continue
                if pred:
                    self.compute_ok()
                for q in range(8):
                    self.ld4(tuple('e%d' % (4 * q + i) for i in range(4)), 4 * q * self.esz)
                if pred:
                    for i in range(32):
                        diag = (i % 2 == 0) and ((i // 2) // 4 == (i // 2) % 4)
                        self.emit('selp.%s e%d, e%d, %s, ok;' % (ty, i, i, self.one() if diag else self.zero()))
                    self.emit('bra.uni B_GEN2_%d;' % k)
                else:
                    self.label('B_GEN2_%d' % k)
                    for xs in groups:
                        T = gen2_temps(1)[0]
                        for y in range(4):
                            gen2_row(y, xs, T, lambda yy, x: ('e%d' % (2 * (4 * yy + x)), 'e%d' % (2 * (4 * yy + x) + 1)))
                        for y in range(4):
                            dr = 'e%d' % (2 * (5 * y))
                            self.emit('fma.rn.%s %s, %s, %s, %s;' % (ty, self.re(xs[y]), dr, self.re(xs[y]), T[y][0]))
                            self.emit('fma.rn.%s %s, %s, %s, %s;' % (ty, self.im(xs[y]), dr, self.im(xs[y]), T[y][1]))
                    self.done_arm()
        gen2_family(False)
        gen2_family(True)
        body = self.lines[body_start:]
        self.lines = self.lines[:body_start]
        return body

    def render(self, fname, cxx_t, cons):
        body = self.build()
        ty = self.ty
        head = []
        head.append('.reg .b32 w0, w1, w2, w3, k0, k1, k2, k3, cls, cf, opa, t0, t1, t2, t3, rc, rv;')
        head.append('.reg .b64 tt;')
        head.append('.reg .pred ok, pt;')
        head.append('.reg .%s c0, c1, c2, c3, c4, c5, c6, c7, n1, n3, n5, n7, sw;' % ty)
        if self.with_gen2:
            head.append('.reg .%s %s;' % (ty, ', '.join('e%d' % i for i in range(32))))
            head.append('.reg .%s ngt;' % ty)
        temps = sorted(self.temps)
        for i in range(0, len(temps), 8):
            head.append('.reg .%s %s;' % (ty, ', '.join(temps[i:i + 8])))
        if self.packed():
            head.append('.reg .b64 Q0, Q1, Q2, Q3, QM2, %s;' % ', '.join(self.pk(x) for x in range(self.nx)))
            t64 = sorted(self.temps64)
            for i in range(0, len(t64), 8):
                head.append('.reg .b64 %s;' % ', '.join(t64[i:i + 8]))
            head.append('mov.b64 QM2, 0xC0000000C0000000;')
        targets = ', '.join(self.labels)
        # software pipeline: the record of op i+1 is fetched while op i runs; the first four coefficients of
        # op i are fetched in the head (every arm but HAD / CX / END uses them) so that their latency overlaps
        # the jump-table load and the indirect branch
        head.append('mov.u32 opa, %%%d;' % self.OPS)
        head.append('ld.shared.v4.u32 {w0, w1, w2, w3}, [opa];')
        head.append('add.u32 opa, opa, 16;')
        head.append('L_NEXT:')
        head.append('and.b32 cls, w0, 0xffff;')
        head.append('add.u32 cf, w2, %%%d;' % self.POOL)
        head.append('ld.shared.v4.u32 {w0, w1, w2, w3}, [opa];')
        head.append('add.u32 opa, opa, 16;')
        if ty == 'f32':
            head.append('ld.shared.v4.f32 {c0, c1, c2, c3}, [cf];')
        else:
            head.append('ld.shared.v2.f64 {c0, c1}, [cf];')
            head.append('ld.shared.v2.f64 {c2, c3}, [cf+16];')
        head.append('TBL: .branchtargets %s;' % targets)
        head.append('brx.idx.uni cls, TBL;')
        tail = ['L_DONE:']
        out = []
        out.append('__device__ __forceinline__ void %s(%s (&a)[%d], %s &ph, uint32_t ops_smem, uint32_t pool_smem,'
                   % (fname, cxx_t, self.nx, cxx_t))
        out.append('                                              uint32_t base, uint64_t outer_ok) {')
        out.append('    asm volatile("{\\n\\t"')
        for l in head + body + tail:
            out.append('        "%s\\n\\t"' % l)
        out.append('        "}"')
        outs = ', '.join('"+%s"(a[%d].x), "+%s"(a[%d].y)' % (cons, x, cons, x) for x in range(self.nx))
        out.append('        : %s, "+%s"(ph.x), "+%s"(ph.y)' % (outs, cons, cons))
        out.append('        : "r"(ops_smem), "r"(pool_smem), "r"(base), "l"(outer_ok)')
        out.append('        : "memory");')
        out.append('}')
        return out, len(self.labels)


def main():
    f32 = Gen('f32', 4)
    d64 = Gen('f64', 3)
    o1, n1 = f32.render('run_sweep_ops_dense2', 'float2', 'f')
    o2, n2 = d64.render('run_sweep_ops_dense2', 'double2', 'd')
    # the lean interpreter: the same arms without the dense two-target family (sweeps that hold no such op use
    # it; with the 4x4 arms in the same block the common arms lost ~3 % on the headline circuit)
    l32 = Gen('f32', 4, with_gen2=False)
    l64 = Gen('f64', 3, with_gen2=False)
    o3, n3 = l32.render('run_sweep_ops', 'float2', 'f')
    o4, n4 = l64.render('run_sweep_ops', 'double2', 'd')
    assert n3 == n1 and n4 == n1
    assert n1 == n2 and f32.labels.index('A_END') == d64.labels.index('A_END')
    names = {}
    for i, l in enumerate(f32.labels):
        names.setdefault(l, i)
    print('// GENERATED by gen_regops.py -- do not edit.  The register-op interpreter of the tile engine (inline PTX).')
    print('// arm ids (RegOp::code); an id + j addresses register bit j, pair ids follow (0,1) (0,2) (0,3) (1,2) (1,3) (2,3)')
    def first(prefix):
        return min(i for l, i in names.items() if l.startswith(prefix))
    consts = [('ARM_HAD', 'A_HAD0'), ('ARM_REAL', 'A_REAL0'), ('ARM_REAL_P', 'A_REALP0'), ('ARM_REALU', 'A_REALU0'),
              ('ARM_IMAG', 'A_IMAG0'), ('ARM_IMAG_P', 'A_IMAGP0'), ('ARM_IMAGU', 'A_IMAGU0'), ('ARM_GEN', 'A_GEN0'),
              ('ARM_GEN_P', 'A_GENP0'), ('ARM_GENM', 'A_GENM0'), ('ARM_CX', 'A_CX0_1'), ('ARM_PH1', 'A_PH10'),
              ('ARM_PH1_P', 'A_PH1P0'), ('ARM_SGN1', 'A_SGN10'), ('ARM_SGN1_P', 'A_SGN1P0'), ('ARM_PH2', 'A_PH20'),
              ('ARM_PH2_P', 'A_PH2P0'), ('ARM_SGN2', 'A_SGN20'), ('ARM_SGN2_P', 'A_SGN2P0'), ('ARM_PHT', 'A_PHT'),
              ('ARM_PHM', 'A_PHM'), ('ARM_END', 'A_END'), ('ARM_GEN2', 'A_GEN20'), ('ARM_GEN2_P', 'A_GEN2P0')]
    # (ids come from the full table; the lean table has the same length)
    for c, l in consts:
        print('constexpr int %s = %d;' % (c, f32.labels.index(l)))
    print('constexpr int ARM_COUNT = %d;' % n1)
    print('__host__ __device__ constexpr int cx_arm(int j, int cb) { return ARM_CX + j * 3 + (cb < j ? cb : cb - 1); }')
    print('constexpr int OUTER_SLOT_NONE = 63;   // bit 63 of the tile-level predicate mask is always 1')
    print()
    for o in (o3, o4, o1, o2):
        print('\n'.join(o))
        print()


if __name__ == '__main__':
    main()
