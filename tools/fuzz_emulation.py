#!/usr/bin/env python3
"""Randomised soak of the device templates compiled for the host (tests/_build/libgb200_hostemu*.so: the default build and
the A/B arithmetic build) against big-int arithmetic and the C++ oracle - longer than the unit tests allow:

   python tools/fuzz_emulation.py field 60000      # mul / sqr / mul_sub / Fp2 mul, sqr on every field, special values mixed in
   python tools/fuzz_emulation.py msm 60           # small MSMs with equal points, P and -P, infinities, special scalars
   python tools/fuzz_emulation.py ba 300           # batched-affine levels on adversarial inputs (few distinct points, +-P)
   python tools/fuzz_emulation.py ntt 400          # NTT pass walks, default and register rounds, random sizes / tiles / modes
   python tools/fuzz_emulation.py fixed 25         # fixed-base batch templates against the C++ oracle (slow: emulated tables)
   python tools/fuzz_emulation.py pipes 12         # FP64-pipe accumulate and the hybrid split
   python tools/fuzz_emulation.py plonk 40         # C++ PLONK orchestration (mocked kernels) on random instances + pairing Verify

Round 1: 1.4 M products + 720 k single-reduction mul_sub + Fp2 operations and 2 080 MSMs (default build, A/B build, and the
persistent / shared-memory-accumulator variants of the accumulate stage), 1 920 batched-affine MSMs, 7 600 NTTs, 156 fixed-base batches, 78 FP64-pipe / hybrid MSMs,
160 PLONK proofs: no mismatch.  (The NTT soak did find a bound that was missing in ntt_make_plan for tiles of 2^3 at 2^13 points - emulation
only, device tiles are 2^6 and up - now guarded.)"""
import sys
MODE = sys.argv[1] if len(sys.argv) > 1 else "field"
sys.argv = [sys.argv[0]] + sys.argv[2:]
if MODE == "field":
    import sys, ctypes, random, time
    import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import numpy as np
    from oracle import ff
    from oracle.params import CURVES
    P=lambda a:a.ctypes.data_as(ctypes.c_void_p)
    libs={'default':ctypes.CDLL(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))+'/tests/_build/libgb200_hostemu.so'),'opt':ctypes.CDLL(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))+'/tests/_build/libgb200_hostemu_opt.so')}
    rng=random.Random(12345)
    N=int(sys.argv[1]) if len(sys.argv)>1 else 5000
    def special(q,L):
        top=1<<(64*L)
        c=[0,1,q-1,q-2,(q+1)//2,2**32-1,2**64-1,(1<<(q.bit_length()-1)),q-(1<<32),q-(1<<64)]
        return [x%q for x in c]
    t0=time.time(); bad=0
    for name,lib in libs.items():
      for c in CURVES.values():
        for which,(q,L) in enumerate(((c.p,c.fp_limbs),(c.r,c.fr_limbs))):
          fid=c.curve_id*2+which
          sp=special(q,L)
          def pick(): return rng.choice(sp) if rng.random()<0.15 else rng.randrange(q)
          for it in range(N):
            a,b,cc,d=pick(),pick(),pick(),pick()
            A=ff.pack_elements([a],q,L); B=ff.pack_elements([b,cc,d],q,L); O=np.zeros_like(A)
            lib.emu_field_op(fid,2,P(A),P(B),P(O)); 
            if ff.unpack_elements(O,q,L)[0]!=a*b%q: bad+=1; print('MUL',name,c.name,which,hex(a),hex(b))
            lib.emu_field_op(fid,5,P(A),P(A),P(O))
            if ff.unpack_elements(O,q,L)[0]!=a*a%q: bad+=1; print('SQR',name,c.name,which,hex(a))
            if which==0:
              lib.emu_field_op(fid,8,P(A),P(B),P(O))
              if ff.unpack_elements(O,q,L)[0]!=(a*b-cc*d)%q: bad+=1; print('MULSUB',name,c.name,hex(a),hex(b),hex(cc),hex(d))
        # Fp2
        for c in [c for c in CURVES.values() if c.fp2_nonresidue is not None]:
          F2=ff.Fp2(c.p,c.fp2_nonresidue); L=c.fp_limbs
          for it in range(N//2):
            a=(rng.randrange(c.p),rng.randrange(c.p)); b=(rng.randrange(c.p),rng.randrange(c.p))
            A=ff.pack_elements(list(a),c.p,L).reshape(-1); B=ff.pack_elements(list(b),c.p,L).reshape(-1); O=np.zeros_like(A)
            lib.emu_field_op(100+c.curve_id*2,2,P(A),P(B),P(O))
            if tuple(ff.unpack_elements(O,c.p,L))!=F2.mul(a,b): bad+=1; print('FP2MUL',name,c.name)
            lib.emu_field_op(100+c.curve_id*2,5,P(A),P(A),P(O))
            if tuple(ff.unpack_elements(O,c.p,L))!=F2.sqr(a): bad+=1; print('FP2SQR',name,c.name)
    print('done',N,'bad',bad,round(time.time()-t0,1),'s')

elif MODE == "msm":
    import sys, ctypes, random, time
    import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import numpy as np
    from oracle import ff, ec, corelib, derive
    from oracle.params import CURVES
    P=lambda a:a.ctypes.data_as(ctypes.c_void_p)
    libs={'default':ctypes.CDLL(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))+'/tests/_build/libgb200_hostemu.so'),'opt':ctypes.CDLL(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))+'/tests/_build/libgb200_hostemu_opt.so')}
    rng=random.Random(777); t0=time.time(); bad=0; runs=0
    REPS=int(sys.argv[1])
    for c in CURVES.values():
      for group in (1,2):
        F=ff.base_field(c,group); base=derive.subgroup_point(c,group)
        deg=F.degree
        pool=corelib.fixed_base(c,group,ec.pack_points(c,group,[base]),ff.pack_elements([rng.randrange(1,c.r) for _ in range(64)],c.r,c.fr_limbs))
        pool=pool.reshape(64,-1)
        for rep in range(REPS if c.fp_limbs<=6 else max(1,REPS//4)):
          n=rng.choice([1,2,5,16,33])
          idx=[rng.randrange(64) for _ in range(n)]
          if n>2 and rng.random()<0.5: idx[1]=idx[0]           # equal points
          pts=np.ascontiguousarray(pool[idx].copy())
          if n>3 and rng.random()<0.3: pts[2]=0                 # infinity
          sc=[rng.choice([0,1,c.r-1,2,rng.randrange(c.r)]) if rng.random()<0.3 else rng.randrange(c.r) for _ in range(n)]
          if n>2 and rng.random()<0.5: sc[1]=sc[0]
          if n>2 and rng.random()<0.3: sc[1]=(c.r-sc[0])%c.r    # P and -P meet
          SA=ff.pack_elements(sc,c.r,c.fr_limbs)
          want=corelib.msm(c,group,pts,SA,c=4)
          wa=ec.from_jac(F,ec.unpack_points(c,group,want,ncoords=3)[0])
          for name,lib in libs.items():
            cw=rng.choice([3,5,8]); pre=rng.choice([0,1]) if c.fp_limbs<=6 else 0
            out=np.zeros(3*deg*c.fp_limbs,dtype=np.uint64)
            rc=lib.emu_msm(c.curve_id,group,P(pts),P(SA),n,cw,pre,rng.choice([2,3,64]),rng.choice([4,16]),P(out))
            ga=ec.from_jac(F,ec.unpack_points(c,group,out,ncoords=3)[0])
            runs+=1
            if rc!=0 or ga!=wa: bad+=1; print('BAD',name,c.name,group,n,cw,pre)
            if name=='default':
              # the run-time accumulate variants (GB200_MSM_PERSISTENT=1/2, GB200_MSM_SMEM_ACC) on the same input
              for tag,call in (('persistent',lambda o: lib.emu_msm_persistent(c.curve_id,group,P(pts),P(SA),n,cw,pre,3,8,rng.choice([1,3,50]),P(o))),
                               ('persistent_smem',lambda o: lib.emu_msm_persistent(c.curve_id,group,P(pts),P(SA),n,cw,pre,3,8,-rng.choice([1,4]),P(o))),
                               ('smem',lambda o: lib.emu_msm_smem(c.curve_id,group,P(pts),P(SA),n,cw,pre,3,8,P(o)))):
                out=np.zeros(3*deg*c.fp_limbs,dtype=np.uint64)
                rc=call(out); runs+=1
                if rc!=0 or ec.from_jac(F,ec.unpack_points(c,group,out,ncoords=3)[0])!=wa: bad+=1; print('BAD',tag,c.name,group,n,cw,pre,rc)
    print('runs',runs,'bad',bad,round(time.time()-t0,1),'s')


if MODE in ("ba", "ntt"):
    # appended modes:  ba <reps>  - batched-affine levels (GB200_MSM_BATCH_AFFINE) on adversarial small MSMs;
    #                  ntt <reps> - NTT pass walks (default and register rounds) at random sizes / tile sizes / modes
    import ctypes, os, random, time
    import numpy as np
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    from oracle import corelib, derive, ec, ff
    from oracle.params import CURVES
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib = ctypes.CDLL(ROOT + "/tests/_build/libgb200_hostemu.so")
    REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    rng = random.Random(4242)
    t0 = time.time(); bad = runs = 0
    if MODE == "ba":
        for c in CURVES.values():
            for group in (1, 2):
                F = ff.base_field(c, group); base = derive.subgroup_point(c, group); deg = F.degree
                pool = corelib.fixed_base(c, group, ec.pack_points(c, group, [base]),
                                          ff.pack_elements([rng.randrange(1, c.r) for _ in range(16)], c.r, c.fr_limbs)).reshape(16, -1)
                for rep in range(REPS if c.fp_limbs <= 6 else max(1, REPS // 5)):
                    n = rng.choice([3, 17, 64, 90])
                    idx = [rng.randrange(rng.choice([2, 4, 16])) for _ in range(n)]          # few distinct points: P + P, 2P + 2P ...
                    pts = np.ascontiguousarray(pool[idx].copy())
                    if rng.random() < 0.3: pts[rng.randrange(n)] = 0
                    small = [rng.randrange(c.r) for _ in range(3)]
                    sc = [rng.choice(small + [0, 1, c.r - 1, (c.r - small[0]) % c.r]) if rng.random() < 0.7 else rng.randrange(c.r) for _ in range(n)]
                    SA = ff.pack_elements(sc, c.r, c.fr_limbs)
                    want = ec.from_jac(F, ec.unpack_points(c, group, corelib.msm(c, group, pts, SA, c=4), ncoords=3)[0])
                    cw = rng.choice([3, 4, 6]); levels = rng.choice([1, 2, 3, 5, 12]); tl = rng.choice([2, 4, 64]); ch = rng.choice([2, 8])
                    out = np.zeros(3 * deg * c.fp_limbs, dtype=np.uint64)
                    rc = lib.emu_msm_ba(c.curve_id, group, P(pts), P(SA), n, cw, 0, tl, ch, levels, P(out)); runs += 1
                    if rc != 0 or ec.from_jac(F, ec.unpack_points(c, group, out, ncoords=3)[0]) != want:
                        bad += 1; print("BAD ba", c.name, group, n, cw, levels, tl, ch, rc)
    else:
        for c in CURVES.values():
            for rep in range(REPS):
                logn = rng.randrange(0, 14); tile = rng.randrange(3, 12); r8 = rng.randrange(2)
                inv, dec, cos = rng.randrange(2), rng.randrange(2), rng.randrange(2)
                n = 1 << logn
                A = ff.pack_elements([rng.randrange(c.r) for _ in range(n)], c.r, c.fr_limbs)
                want = corelib.ntt(c, A.copy(), logn, bool(inv), dec, bool(cos))
                lib.emu_ntt_set_tile_log(max(2, tile)); lib.emu_ntt_set_radix8(r8)
                B = A.copy()
                rc = lib.emu_ntt(c.curve_id, P(B), logn, inv, dec, cos, None, None); runs += 1
                if rc != 0 or not np.array_equal(B, want):
                    bad += 1; print("BAD ntt", c.name, logn, tile, r8, inv, dec, cos, rc)
        lib.emu_ntt_set_tile_log(11); lib.emu_ntt_set_radix8(0)
    print("runs", runs, "bad", bad, round(time.time() - t0, 1), "s")

if MODE in ("fixed", "pipes"):
    # fixed <reps> - b200_fixed_base_batch's templates against the C++ oracle's fixed-base batch (random bases, window sizes,
    #                batch sizes around the 16-point inversion chunks, special scalars);
    # pipes <reps> - the FP64-pipe accumulate (emu_msm52) and the hybrid split (emu_msm_hybrid) on random precomputed-table MSMs
    import ctypes, os, random, time
    import numpy as np
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    from oracle import corelib, derive, ec, ff
    from oracle.params import CURVES
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib = ctypes.CDLL(ROOT + "/tests/_build/libgb200_hostemu.so")
    REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    rng = random.Random(99)
    t0 = time.time(); bad = runs = 0
    for c in CURVES.values():
        for group in ((1, 2) if MODE == "fixed" else (1,)):
            if group == 2 and c.fp2_nonresidue is None:
                continue
            F = ff.base_field(c, group); gen = derive.subgroup_point(c, group); deg = F.degree
            for rep in range(REPS if c.fp_limbs <= 6 else max(1, REPS // 4)):
                if MODE == "fixed":
                    base = ec.scalar_mul(F, rng.randrange(1, c.r), gen)
                    n = rng.choice([1, 15, 16, 17, 40])
                    ks = [rng.choice([0, 1, 2, c.r - 1, c.r - 2, (1 << (c.r.bit_length() - 1)) - 1]) if rng.random() < 0.3 else rng.randrange(c.r) for _ in range(n)]
                    KS = ff.pack_elements(ks, c.r, c.fr_limbs); BA = ec.pack_points(c, group, [base])
                    want = corelib.fixed_base(c, group, BA, KS)
                    out = np.zeros((n, 2 * deg * c.fp_limbs), dtype=np.uint64)
                    rc = lib.emu_fixed_base(c.curve_id, group, P(BA), P(KS), n, rng.choice([2, 3, 5, 8, 11]), P(out)); runs += 1
                    if rc != 0 or not np.array_equal(out.reshape(-1), np.asarray(want).reshape(-1)):
                        bad += 1; print("BAD fixed", c.name, group, n, rc)
                else:
                    n = rng.choice([5, 37, 70])
                    pts = corelib.fixed_base(c, 1, ec.pack_points(c, 1, [gen]), ff.pack_elements([rng.randrange(1, c.r) for _ in range(n)], c.r, c.fr_limbs))
                    pts = np.ascontiguousarray(np.asarray(pts).reshape(n, -1))
                    if n > 6: pts[3] = 0; pts[5] = pts[4]
                    sc = [rng.choice([0, 1, c.r - 1]) if rng.random() < 0.2 else rng.randrange(c.r) for _ in range(n)]
                    if n > 6: sc[4] = sc[5]
                    SA = ff.pack_elements(sc, c.r, c.fr_limbs)
                    want = ec.from_jac(F, ec.unpack_points(c, 1, corelib.msm(c, 1, pts, SA, c=4), ncoords=3)[0])
                    cw = rng.choice([4, 7]) if c.fp_limbs > 6 else rng.choice([4, 7, 10])
                    for tag, call in (("msm52", lambda o: lib.emu_msm52(c.curve_id, P(pts), P(SA), n, cw, rng.choice([2, 5, 64]), rng.choice([4, 16]), P(o))),
                                      ("hybrid", lambda o: lib.emu_msm_hybrid(c.curve_id, P(pts), P(SA), n, cw, rng.choice([2, 5]), rng.choice([4, 16]), rng.randrange(1, 16), P(o)))):
                        out = np.zeros(3 * c.fp_limbs, dtype=np.uint64)
                        rc = call(out); runs += 1
                        if rc != 0 or ec.from_jac(F, ec.unpack_points(c, 1, out, ncoords=3)[0]) != want:
                            bad += 1; print("BAD", tag, c.name, n, cw, rc)
    print("runs", runs, "bad", bad, round(time.time() - t0, 1), "s")

if MODE == "plonk":
    # plonk <reps> - the C++ PLONK orchestration on the mocked C ABI, random instances (sizes, BSB22 gates, cache on/off, two
    # curves) against the oracle prover, every third one also through the pairing verifier
    import sys, os, ctypes, random, time
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + '/tests')
    import numpy as np
    from gnark_b200 import lib as b200
    from oracle import corelib, ec, ff, plonk_prover as pp
    from oracle.params import CURVES
    from util import jac_to_affine
    m = ctypes.CDLL(ROOT + '/tests/_build/libgb200_plonkmock.so')
    vp, i32 = ctypes.c_void_p, ctypes.c_int32
    m.b200_plonk_pk_load.argtypes = [i32, i32, ctypes.POINTER(b200.PlonkPkDesc), ctypes.POINTER(vp)]
    m.b200_plonk_pk_free.argtypes = [vp]
    m.b200_plonk_prove.argtypes = [vp, vp, vp, vp, ctypes.POINTER(b200.PlonkChallenges), vp, vp]
    m.b200_last_error.restype = ctypes.c_char_p
    rng = random.Random(2026); bad = 0; t0 = time.time(); runs = 0
    REPS = int(sys.argv[1])
    for rep in range(REPS):
        c = CURVES[rng.choice(["bn254", "bls12-381"])]
        logn = rng.choice([3, 4, 5, 6]); n = 1 << logn; r, L = c.r, c.fr_limbs
        ncom = rng.choice([0, 0, 1, 2])
        inst = pp.random_satisfied_instance(c, n, seed=rng.randrange(1 << 30), n_commit=ncom)
        circ, l, rr, o = inst[:4]; pi2 = inst[4] if ncom else ()
        rnd = lambda: rng.randrange(r)
        ch = pp.Challenges(gamma=rnd(), beta=rnd(), alpha=rnd(), zeta=rnd(), v=rnd(), bl=[rnd(), rnd()], br=[rnd(), rnd()], bo=[rnd(), rnd()], bz=[rnd(), rnd(), rnd()])
        tau = rnd()
        want = pp.prove(c, circ, l, rr, o, ch, tau, pi2=pi2) if ncom else pp.prove(c, circ, l, rr, o, ch, tau)
        pe = lambda v: np.ascontiguousarray(ff.pack_elements(v, r, L)); P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        srs = np.ascontiguousarray(corelib.fixed_base(c, 1, ec.pack_points(c, 1, [c.g1]), pe([pow(tau, i, r) for i in range(n + 3)])))
        keep = {k: pe(getattr(circ, k)) for k in ("ql", "qr", "qm", "qo", "qk")}
        perm = np.ascontiguousarray(np.array(circ.perm, dtype=np.int64))
        d = b200.PlonkPkDesc(); d.log2n = logn
        for k, a in keep.items(): setattr(d, k, P(a).value)
        d.perm, d.srs_canonical = P(perm).value, P(srs).value
        qcp = [pe(v) for v in circ.qcp]
        if ncom:
            qarr = (ctypes.c_void_p * ncom)(*[a.ctypes.data for a in qcp]); d.n_qcp, d.qcp = ncom, ctypes.cast(qarr, ctypes.POINTER(ctypes.c_void_p))
        os.environ["GB200_PLONK_COSET_CACHE"] = rng.choice(["0", "1"])
        h = ctypes.c_void_p(0)
        assert m.b200_plonk_pk_load(0, c.curve_id, ctypes.byref(d), ctypes.byref(h)) == 0, m.b200_last_error()
        sc = {k: pe(v) for k, v in (("gamma", [ch.gamma]), ("beta", [ch.beta]), ("alpha", [ch.alpha]), ("zeta", [ch.zeta]), ("v", [ch.v]), ("bl", ch.bl), ("br", ch.br), ("bo", ch.bo), ("bz", ch.bz))}
        cs = b200.PlonkChallenges()
        for k, a in sc.items(): setattr(cs, k, P(a).value)
        pts = np.zeros((10, 3 * c.fp_limbs), dtype=np.uint64); vals = np.zeros((7 + ncom, L), dtype=np.uint64); bsb = np.zeros((max(ncom,1), 3 * c.fp_limbs), dtype=np.uint64)
        L_, R_, O_ = pe(l), pe(rr), pe(o)
        if ncom:
            pi2a = [pe(v) for v in pi2]; parr = (ctypes.c_void_p * ncom)(*[a.ctypes.data for a in pi2a])
            cs.pi2, cs.out_bsb22 = ctypes.cast(parr, ctypes.POINTER(ctypes.c_void_p)), P(bsb).value
        rc = m.b200_plonk_prove(h, P(L_), P(R_), P(O_), ctypes.byref(cs), P(pts), P(vals)); runs += 1
        F = ff.Fp(c.p)
        dl = [want.L, want.R, want.O, want.Z, want.H[0], want.H[1], want.H[2], want.lin, want.batch_opening, want.z_opening]
        ok = rc == 0 and all(jac_to_affine(c, 1, pts[k]) == ec.scalar_mul(F, dl[k], c.g1) for k in range(10))
        got = ff.unpack_elements(vals, r, L)
        ok = ok and got[:6] + got[7:] == want.claimed and got[6] == want.zu
        if ok and rep % 3 == 0:
            ok = pp.verify_pairing(c, circ, [jac_to_affine(c, 1, pts[k]) for k in range(10)], got, ch, tau, bsb22_points=[jac_to_affine(c, 1, bsb[j]) for j in range(ncom)])
        if not ok: bad += 1; print("BAD", c.name, logn, ncom, rc, m.b200_last_error())
        m.b200_plonk_pk_free(h)
    print("runs", runs, "bad", bad, round(time.time() - t0, 1), "s")

