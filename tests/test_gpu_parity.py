"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI (ctypes -> libgjkepa_b200.so),
against the oracle on the same seeded inputs, the committed golden fixture, and size-independent
properties at full size.  Bars: collision flags and GJK exit path / walk iterations BIT-EXACT;
penetration depth |d_gpu - d_ref| <= 1e-9 * max(d_ref, 1) (unit hull scale) and normal
1 - |cos| <= 1e-9 (north_star: "about 1e-9 in real(8)")."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DEPTH_RTOL = 1e-9
NORMAL_TOL = 1e-9


def configure(eng, variant):
    """Kernel variants that must all give the same bits: 0 = automatic (init kernel + compacting walk kernel,
    re-pack threshold 8; free-running 4-lane init kernel for larger hulls; linked-mesh EPA kernels first), "r1"/"r32" = other
    re-pack thresholds, "mono" = one GJK kernel per tile, "gen" = the general EPA kernel for every contact, 1/2/4 = sub-warp GJK kernels with that many lanes per pair."""
    eng.set_tuning(eng.TUNE_SINGLE_KERNEL, 1 if variant == "mono" else 0)
    eng.set_tuning(eng.TUNE_EPA_GENERAL, 1 if variant == "gen" else 0)
    eng.set_tuning(eng.TUNE_WALK_REFILL, int(variant[1:]) if isinstance(variant, str) and variant[0] == "r" else 8)
    eng.set_tuning(eng.TUNE_EARLY_OUT, 0)                  # the parity build; test_early_out_* switches it on
    eng.set_lanes(variant if isinstance(variant, int) else 0)


ALL_VARIANTS = (0, "r1", "r32", "mono", "gen", 1, 2, 4)


def check_against_oracle(orc, eng, b, mode, lanes=ALL_VARIANTS, version=1, tol_ff=1.0):
    omode = {"gjk": orc.MODE_GJK, "epa": orc.MODE_EPA, "full": orc.MODE_FULL}[mode]
    ref = orc.batch(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=omode, version=version, tol_ff=tol_ff)
    for L in lanes:
        configure(eng, L)
        got = eng.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=mode, version=version, tol_ff=tol_ff)
        assert np.array_equal(got["collision"], ref["collision"]), f"flags differ (lanes={L})"
        assert np.array_equal(got["status"] & 0xffff, ref["status"] & 0xffff), f"GJK exit path/iterations differ (lanes={L})"
        if mode == "gjk":
            continue
        hit = ref["collision"] > 0
        clean = hit & ((ref["status"] >> 24) == 0)
        if not clean.any():
            assert np.all(got["depth"][~hit] == 0.0)
            continue
        assert np.all((got["status"][clean] >> 24) == 0)
        d_ref, d_got = ref["depth"][clean], got["depth"][clean]
        assert np.all(np.abs(d_got - d_ref) <= DEPTH_RTOL * np.maximum(d_ref, 1.0))
        cosd = np.sum(got["normal"][clean] * ref["normal"][clean], axis=1)
        if version == 1:
            assert np.all(1.0 - cosd <= NORMAL_TOL)                  # same orientation too
        # non-colliding pairs keep the zeroed outputs of F:65-71
        assert np.all(got["depth"][~hit] == 0.0) and np.all(got["normal"][~hit] == 0.0)
        if mode == "full":
            assert np.array_equal(got["colli_type"][clean], ref["colli_type"][clean])
            assert np.all(got["colli_type"][~hit] == 0) and np.all(got["nearest"][~hit] == 0.0)
            # The epilogue's outputs are DISCONTINUOUS functions of the normal (a selected vertex, a mean over the vertices
            # within a band of the maximum: F:722-747, F:766-804), so they are not compared with the oracle's values --
            # whose normal may differ in the last bit -- but with the reference's own rules (oracle, F:326-343) evaluated
            # ON THE NORMAL THE GPU RETURNED: bit for bit, every row, no tolerance and no allowed fraction of misses.
            nrm_in = got["normal"]
            if version == 3:             # F:335 overwrites collision_normal_; the normal EPA found comes from an EPA-mode call
                nrm_in = eng.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="epa", version=version, tol_ff=tol_ff)["normal"]
            rows = np.nonzero(clean)[0]
            if len(rows) > 6000:
                rows = rows[np.random.default_rng(5).choice(len(rows), 6000, replace=False)]
            bad = []
            for i in rows:
                ep = orc.epilogue(b.hull(b.pair_a[i]), b.hull(b.pair_b[i]), nrm_in[i], version=version, tol_ff=tol_ff)
                ok = (np.array_equal(ep["point"], got["point"][i], equal_nan=True) and ep["colli_type"] == got["colli_type"][i]
                      and np.array_equal(ep["nearest"], got["nearest"][i]) and np.array_equal(ep["normal"], got["normal"][i], equal_nan=True))
                if not ok:
                    bad.append((int(i), ep["point"].tolist(), got["point"][i].tolist(), ep["colli_type"], int(got["colli_type"][i])))
            assert not bad, f"epilogue differs from the reference rule on the returned normal (lanes={L}): {len(bad)} of {len(rows)} rows, e.g. {bad[:3]}"
            # and against the oracle's own rows wherever both normals are the same bits
            same_n = clean & np.all(got["normal"] == ref["normal"], axis=1)
            assert np.array_equal(got["point"][same_n], ref["point"][same_n], equal_nan=True)
    return ref


@pytest.mark.parametrize("off,coll,depth,nrm", [
    ((0.5, 0.0, 0.0), True, 0.5, (1, 0, 0)),
    ((0.5, 0.25, 0.125), True, 0.5, (1, 0, 0)),
    ((1.5, 0.0, 0.0), False, 0.0, (0, 0, 0)),
    ((1.0, 0.0, 0.0), True, 0.0, (1, 0, 0)),
])
def test_cfg1_unit_cubes_through_the_drop_in_call(W, orc, off, coll, depth, nrm):
    # BASELINE.json configs[0], through gjkepa_pair_f64 (all scalars by reference)
    from collision_detect_gjk_epa_b200 import GJKEPA
    b = W.unit_cubes(off)
    col, typ, near, n, pt, dep = GJKEPA(1, 1.0, b.hull(0), b.hull(1))
    assert col == coll and abs(dep - depth) < 1e-12 and np.allclose(n, nrm, atol=1e-12)
    o = orc.pair(b.hull(0), b.hull(1))
    assert col == o["collision"] and typ == o["colli_type"] and np.allclose(pt, o["point"], atol=1e-12)
    assert near.shape == (2, 3)
    if coll:
        assert abs(near[0] @ n - 0.5) < 1e-12                        # p1 support along +normal


def test_cfg1_cubes_batch_all_lane_widths(W, orc, engine):
    for off in [(0.5, 0, 0), (0.5, 0.25, 0.125), (1.5, 0, 0), (1.0, 0, 0), (0.5, 0.5, 0), (0.25, 0.5, 0.75), (1, 1, 1)]:
        check_against_oracle(orc, engine, W.unit_cubes(off), "full")


def test_coincident_cubes_defined_status(W, orc, engine):
    b = W.unit_cubes((0.0, 0.0, 0.0))                                # hazard H1: never blocks, never exits
    got = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="full")
    assert got["collision"][0] == 1 and (got["status"][0] & 0xff) == 4 and got["status"][0] & (1 << 28)
    assert got["depth"][0] == 0.0


def test_cfg2_box8_gjk_flags_bit_exact(W, orc, engine):
    check_against_oracle(orc, engine, W.make_pairs("box8", 300000, 1.0), "gjk")


def test_box8_full_outputs(W, orc, engine):
    check_against_oracle(orc, engine, W.make_pairs("box8", 60000, 1.0, seed=3), "full")


def test_cfg3_hull32_gjk_epa(W, orc, engine):
    check_against_oracle(orc, engine, W.make_pairs("hull", 120000, 0.85, n_vert=32), "full")


def test_cfg4_mixed_sizes_and_deep_penetration(W, orc, engine):
    check_against_oracle(orc, engine, W.make_mixed(60000, 0.25), "full")
    check_against_oracle(orc, engine, W.make_pairs("hull", 8000, 0.05, n_vert=64), "epa")


def test_cfg5_shared_hull_store(W, orc, engine):
    b = W.make_shared_store(50000, 200000, n_vert=16)
    ref = check_against_oracle(orc, engine, b, "epa", lanes=(0, "r1", "mono", 4))
    assert 0.02 < ref["collision"].mean() < 0.98


def test_small_and_ragged_hulls(W, orc, engine):
    check_against_oracle(orc, engine, W.make_pairs("hull", 100000, 0.15, n_vert=5), "full")
    check_against_oracle(orc, engine, W.make_pairs("hull", 100000, 0.5, n_vert=4), "full")
    check_against_oracle(orc, engine, W.make_mixed(30000, 0.6, seed=5, n_lo=4, n_hi=9), "full")


def test_version3_turntable_contact_point(W, orc, engine):
    check_against_oracle(orc, engine, W.make_pairs("hull", 20000, 0.5, seed=8, n_vert=12), "full", version=3)


def test_version2_general_contact_point(W, orc, engine):
    """version_ = 2 (get_collisionPoint_02, F:457-696): vertex/vertex, vertex/face, edge/edge (FOOT_LL), edge/face
    (SORT_CLOCK + IS_INSIDE_PF + FOOT_PL) and face/face branches, on boxes (faces and edges in contact), small hulls
    and the planted axis-aligned cubes."""
    check_against_oracle(orc, engine, W.make_pairs("box8", 40000, 0.8, seed=17), "full", version=2, lanes=(0, "gen", 4))
    check_against_oracle(orc, engine, W.make_pairs("hull", 30000, 0.4, seed=18, n_vert=10), "full", version=2, lanes=(0, "gen"))
    check_against_oracle(orc, engine, W.make_pairs("hull", 10000, 0.7, seed=19, n_vert=40), "full", version=2, lanes=(0,))


def test_tolff_threshold_is_honoured(W, orc, engine):
    check_against_oracle(orc, engine, W.make_pairs("box8", 20000, 0.8, seed=9), "full", tol_ff=0.05)


def test_golden_fixture(W, golden, engine):
    g = golden
    for L in ALL_VARIANTS:
        configure(engine, L)
        got = engine.batch_host(g["verts"], g["hull_off"], g["pair_a"], g["pair_b"], mode="epa")
        assert np.array_equal(got["collision"], g["collision"])
        assert np.array_equal(got["status"] & 0xff, g["exit_path"])
        assert np.array_equal((got["status"] >> 8) & 0xff, g["gjk_iters"])
        m = g["epa_ok"] > 0
        assert np.all(np.abs(got["depth"][m] - g["depth"][m]) <= DEPTH_RTOL * np.maximum(g["depth"][m], 1.0))
        cosd = np.sum(got["normal"][m] * g["normal"][m], axis=1)
        deep = g["depth"][m] > 1e-9                      # at zero depth (touching) the sign of the normal is
        assert np.all(1.0 - cosd[deep] <= NORMAL_TOL)    # decided by the hull's face winding: compare |cos|
        assert np.all(1.0 - np.abs(cosd[~deep]) <= NORMAL_TOL)


def test_empty_batch_and_single_pair(W, engine):
    b = W.make_pairs("box8", 4, 1.0)
    got = engine.batch_host(b.verts, b.hull_off, b.pair_a[:0], b.pair_b[:0], mode="full", want_contacts=True)
    assert got["collision"].shape == (0,) and got["n_contacts"] == 0
    got = engine.batch_host(b.verts, b.hull_off, b.pair_a[:1], b.pair_b[:1], mode="full", want_contacts=True)
    assert got["collision"].shape == (1,) and got["n_contacts"] == int(got["collision"][0])


def test_contact_list_is_sorted_and_complete(W, engine):
    b = W.make_pairs("hull", 50000, 0.85, n_vert=16)
    got = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="epa", want_contacts=True)
    idx = got["contact_index"][:got["n_contacts"]]
    assert np.array_equal(idx, np.nonzero(got["collision"])[0])


def test_device_resident_entry_matches_host_entry(W, engine):
    import torch
    b = W.make_pairs("hull", 100000, 0.85, n_vert=32)
    host = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="full")
    dev = torch.device("cuda:0")
    t = [torch.from_numpy(x).to(dev) for x in (b.verts, b.hull_off, b.pair_a, b.pair_b)]
    out = engine.batch_device(*t, mode="full", want_contacts=True)
    torch.cuda.synchronize()
    for k in ("collision", "depth", "normal", "colli_type", "nearest", "point"):
        assert np.array_equal(out[k].cpu().numpy(), host[k]), k
    assert np.array_equal(out["status"].cpu().numpy().view(np.uint32), host["status"])
    nc = int(out["n_contacts"].item())
    assert nc == int(host["collision"].sum())
    assert np.array_equal(np.sort(out["contact_index"][:nc].cpu().numpy()), np.nonzero(host["collision"])[0])
    engine.sync()


def test_full_size_properties_1m_pairs(W, orc, engine):
    """BASELINE.json sizes (1M pairs), through properties that need no full oracle run:
    determinism across runs and lane widths, zeroed outputs for misses, unit normals, and
    'translate p2 by (depth +- margin) * normal' separates / keeps the contact."""
    b = W.make_pairs("hull", 1_000_000, 0.85, n_vert=32)
    configure(engine, 0)
    r4 = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="epa")
    engine.set_lanes(1)
    r1 = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="epa")
    for k in ("collision", "status", "depth", "normal"):
        assert np.array_equal(r4[k], r1[k]), k                       # idempotent, lane-width independent
    hit = r4["collision"] > 0
    assert 0.45 < hit.mean() < 0.60                                  # ~52.5 % collide
    assert np.all((r4["status"][hit] >> 24) == 0)
    assert np.all(np.abs(np.linalg.norm(r4["normal"][hit], axis=1) - 1.0) < 1e-12)
    assert np.all(r4["depth"][hit] >= 0.0) and np.all(r4["depth"][~hit] == 0.0)
    # spot-check a bounded sample against the oracle at this size
    sel = np.random.default_rng(0).choice(b.n_pairs, 20000, replace=False).astype(np.int64)
    ref = orc.batch(b.verts, b.hull_off, b.pair_a[sel], b.pair_b[sel], mode=orc.MODE_EPA)
    assert np.array_equal(ref["collision"], r4["collision"][sel])
    h = ref["collision"] > 0
    assert np.all(np.abs(ref["depth"][h] - r4["depth"][sel][h]) <= DEPTH_RTOL * np.maximum(ref["depth"][h], 1.0))
    # separation property on the colliding pairs: shift body B and re-run GJK only
    for sign, expect in ((+1.0, 0), (-1.0, 1)):
        verts = b.verts.copy()
        n = 32
        v = verts.reshape(-1, 3, n)                                  # per hull [x|y|z]
        shift = np.zeros((b.n_pairs, 3))
        shift[hit] = (r4["depth"][hit] + sign * 1e-6)[:, None] * r4["normal"][hit]
        v[1::2] += shift[:, :, None]
        g = engine.batch_host(verts, b.hull_off, b.pair_a, b.pair_b, mode="gjk")
        assert np.all(g["collision"][hit] == expect)


def test_max_hull_hint_and_overflow_to_subwarp_kernel(W, orc, engine):
    """lanes=0 (automatic): hulls up to 16 vertices run one thread per pair over shared-memory columns,
    larger ones overflow to the 4-lane kernel; a caller-supplied bound shrinks the columns.  Results
    must not depend on any of that."""
    configure(engine, 0)
    b = W.make_mixed(40000, 0.4, seed=21, n_lo=4, n_hi=40)           # both paths in one batch
    ref = orc.batch(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=orc.MODE_EPA)
    for hint in (0, 40, 64):
        engine.set_max_hull_vertices(hint)
        got = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="epa")
        assert np.array_equal(got["collision"], ref["collision"])
        assert np.array_equal(got["status"] & 0xffff, ref["status"] & 0xffff)
    b8 = W.make_pairs("box8", 50000, 1.0, seed=22)
    ref8 = orc.batch(b8.verts, b8.hull_off, b8.pair_a, b8.pair_b, mode=orc.MODE_GJK)
    for hint in (8, 12, 0):
        engine.set_max_hull_vertices(hint)
        got = engine.batch_host(b8.verts, b8.hull_off, b8.pair_a, b8.pair_b, mode="gjk")
        assert np.array_equal(got["collision"], ref8["collision"])
        assert np.array_equal(got["status"] & 0xffff, ref8["status"] & 0xffff)
    # a bound the data violates: pairs with a larger hull are NOT processed and say so (GJKEPA_ST_BAD_HINT), the others
    # keep their results -- never a silent miss
    BAD_HINT = np.uint32(1 << 31)
    sizes = np.diff(b.hull_off)
    for hint, limit in ((12, 12), (16, 16), (30, 30)):
        engine.set_max_hull_vertices(hint)
        got = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="epa")
        over = (sizes[b.pair_a] > limit) | (sizes[b.pair_b] > limit)
        flagged = (got["status"] & BAD_HINT) != 0
        assert not flagged[~over].any() and not got["collision"][flagged].any() and np.all(got["depth"][flagged] == 0.0)
        ok = ~flagged
        assert np.array_equal(got["collision"][ok], ref["collision"][ok])
        assert np.array_equal(got["status"][ok] & 0xffff, ref["status"][ok] & 0xffff)
        assert 0 < flagged.sum() <= over.sum()          # a pair over the bound that still fits its column is simply processed
    engine.set_max_hull_vertices(0)


def test_full_size_cfg2_kernel_variants_agree(W, orc, engine):
    """BASELINE.json configs[1] at full size (1M box pairs, flags only): the compacting two-stage path, every
    re-pack threshold, the single-kernel form and the 4-lane sub-warp kernel give identical flags and status
    words; a bounded sample is checked against the oracle."""
    b = W.make_pairs("box8", 1_000_000, 1.0)
    engine.set_max_hull_vertices(8)
    outs = {}
    for v in (0, "r1", "r4", "r16", "r32", "mono", 4):
        configure(engine, v)
        outs[v] = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="gjk", want_contacts=True)
    engine.set_max_hull_vertices(0)
    configure(engine, 0)
    for v, o in outs.items():
        assert np.array_equal(o["collision"], outs[0]["collision"]), v
        assert np.array_equal(o["status"], outs[0]["status"]), v
        assert o["n_contacts"] == int(outs[0]["collision"].sum())
    sel = np.random.default_rng(1).choice(b.n_pairs, 50000, replace=False).astype(np.int64)
    ref = orc.batch(b.verts, b.hull_off, b.pair_a[sel], b.pair_b[sel], mode=orc.MODE_GJK)
    assert np.array_equal(ref["collision"], outs[0]["collision"][sel])
    assert np.array_equal(ref["status"] & 0xffff, outs[0]["status"][sel] & 0xffff)


def test_touching_cubes_inside_lockstep_warps(W, orc, engine):
    """Axis-aligned touching / nearly touching bodies (the reference's 1e-8 bands: on-face containment, coplanar
    exits, origin-on-face EPA iterations) scattered among ordinary pairs, with 20-vertex hulls so that they run in
    the kernels that keep several pairs per warp in lockstep.  Regression: a short-circuited vote once let the
    groups of a warp take different numbers of collectives.  Flags, exit paths and depths must match the oracle
    everywhere; normals are compared on the ordinary pairs only (for the planted cubes several faces tie exactly,
    and at zero depth the sign follows the face winding -- DESIGN.md, known divergences)."""
    b = W.make_touching_mix(40000)
    ref = orc.batch(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=orc.MODE_EPA)
    planted = np.zeros(b.n_pairs, bool); planted[::5] = True
    assert len(set((ref["status"][planted] & 0xff).tolist())) >= 2 and ref["collision"][planted].min() == 0
    for hint in (20, 0):
        engine.set_max_hull_vertices(hint)
        for v in (0, "r1", "r32", "mono", "gen", 4):
            configure(engine, v)
            got = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="epa")
            assert np.array_equal(got["collision"], ref["collision"]), (hint, v)
            assert np.array_equal(got["status"] & 0xffff, ref["status"] & 0xffff), (hint, v)
            hit = (ref["collision"] > 0) & ((ref["status"] >> 24) == 0)
            assert np.all((got["status"][hit] >> 24) == 0), (hint, v)
            assert np.all(np.abs(got["depth"][hit] - ref["depth"][hit]) <= DEPTH_RTOL * np.maximum(ref["depth"][hit], 1.0)), (hint, v)
            m = hit & ~planted
            assert np.all(1.0 - np.sum(got["normal"][m] * ref["normal"][m], axis=1) <= NORMAL_TOL), (hint, v)
            pn = got["normal"][hit & planted]
            assert np.all(np.abs(np.linalg.norm(pn, axis=1) - 1.0) < 1e-12)
    engine.set_max_hull_vertices(0)
    configure(engine, 0)


def test_large_batch_equals_its_chunks(W, engine):
    """cfg5-shaped batch well beyond the other tests (8M pairs over a shared store of 1M 16-vertex bodies): the
    queues between the kernels, the work lists and every 32-bit index must hold up.  The whole batch and the same
    pairs run as four chunks must give identical flags, status words, depths and normals; a second shape with
    24-vertex hulls takes the 4-lane kernels."""
    for n_vert, n_pairs in ((16, 8_000_000), (24, 3_000_000)):
        b = W.make_shared_store(1_000_000, n_pairs, n_vert=n_vert, seed=31)
        configure(engine, 0)
        engine.set_max_hull_vertices(n_vert)
        whole = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode="epa", want_contacts=True)
        assert whole["n_contacts"] == int(whole["collision"].sum()) and 0.05 < whole["collision"].mean() < 0.95
        q = n_pairs // 4
        for c in range(4):
            sl = slice(c * q, (c + 1) * q)
            part = engine.batch_host(b.verts, b.hull_off, b.pair_a[sl], b.pair_b[sl], mode="epa")
            for k in ("collision", "status", "depth", "normal"):
                assert np.array_equal(part[k], whole[k][sl]), (n_vert, c, k)
    engine.set_max_hull_vertices(0)


def test_resident_shapes_and_poses(W, orc, engine):
    """gjkepa_shapes_upload + gjkepa_batch_posed_f64: the pose is applied on the device with the arithmetic stated in
    include/gjkepa.h.  The result must be BIT-IDENTICAL to transforming on the host the same way (oracle/orc.py
    apply_pose) and calling the world-space entry, and must match the oracle on those world-space hulls."""
    configure(engine, 0)
    for kw, mode in ((dict(n_shapes=500, n_bodies=60000, n_pairs=80000, kind="box"), "full"),
                     (dict(n_shapes=300, n_bodies=40000, n_pairs=50000, kind="hull", n_lo=4, n_hi=40), "epa")):
        pb = W.make_posed(**kw)
        engine.upload_shapes(pb.shape_verts, pb.shape_off)
        got = engine.batch_posed(pb.body_shape, pb.pose, pb.pair_a, pb.pair_b, mode=mode, want_contacts=True)
        verts, hull_off = orc.apply_pose(pb.shape_verts, pb.shape_off, pb.body_shape, pb.pose)
        engine.set_max_hull_vertices(int(np.diff(pb.shape_off).max()))   # the posed entry derives this bound from the shapes
        same = engine.batch_host(verts, hull_off, pb.pair_a, pb.pair_b, mode=mode, want_contacts=True)
        engine.set_max_hull_vertices(0)
        nc = same["n_contacts"]
        assert got["n_contacts"] == nc and np.array_equal(got["contact_index"][:nc], same["contact_index"][:nc])
        for k in same:
            if k not in ("n_contacts", "contact_index"):
                assert np.array_equal(got[k], same[k]), k
        assert 0.1 < got["collision"].mean() < 0.9
        b = W.PairBatch(verts, hull_off, pb.pair_a, pb.pair_b)
        check_against_oracle(orc, engine, b, mode, lanes=(0,))
    with pytest.raises(Exception):
        engine.batch_posed(pb.body_shape, pb.pose[:-1], pb.pair_a, pb.pair_b)


def test_large_hulls_take_the_general_kernels(W, orc, engine):
    """Hulls beyond every staging buffer (100 and 200 vertices; 65 + 64 straddles the 128-vertex EPA buffer): the
    4-lane kernel reading global memory, the general EPA kernel with unstaged hulls and the contact kernel's scans."""
    for n_vert, n_pairs, sigma in ((100, 6000, 0.8), (200, 2500, 0.9), (65, 6000, 0.8)):
        b = W.make_pairs("hull", n_pairs, sigma, seed=50 + n_vert, n_vert=n_vert)
        for version in (1, 2):
            ref = check_against_oracle(orc, engine, b, "full", lanes=(0, "gen"), version=version)
        assert 0.2 < ref["collision"].mean() < 0.98
    b = W.make_mixed(8000, 0.4, seed=61, n_lo=50, n_hi=150)
    check_against_oracle(orc, engine, b, "full", lanes=(0, 4))


def test_degenerate_hulls_points_segments_flat_and_duplicated(W, orc, engine):
    """Inputs the reference accepts (p(n,3) assumed-shape, any n >= 1) but a hull library would reject: 1-3 vertex
    'hulls' (point, segment, triangle), planar polygons, repeated vertices.  They leave through the initial-tetrahedron
    exits (F:86 direction table exhausted, F:123 duplicate third point, F:157 coplanar fourth point) far more often than
    random hulls do; flags and exit paths must still be the oracle's, bit for bit, in every kernel variant."""
    b = W.make_mixed(40000, 0.5, seed=3, n_lo=1, n_hi=5)
    ref = check_against_oracle(orc, engine, b, "epa")
    seen = set(np.unique(ref["status"] & 0xff).tolist())
    assert {1, 2, 3, 5, 6, 9, 10} <= seen, seen                       # sphere, init_cap, dup3, coplanar4, contain, coplanar, contain
    flat = W.make_pairs("hull", 30000, 0.5, seed=4, n_vert=8)
    v = flat.verts.reshape(-1, 3, 8)
    v[0::2, 2, :] = v[0::2, 2, :1]                                    # body A: a planar polygon
    check_against_oracle(orc, engine, flat, "epa")
    v[1::2, 2, :] = v[1::2, 2, :1]                                    # both planar (parallel planes): coplanar4 for ~all
    ref = check_against_oracle(orc, engine, flat, "gjk")
    assert np.mean((ref["status"] & 0xff) == 5) > 0.9
    dup = W.make_pairs("hull", 30000, 0.6, seed=5, n_vert=10)
    v = dup.verts.reshape(-1, 3, 10)
    v[:, :, 5:] = v[:, :, :5]                                         # every vertex twice: ties go to the first index
    check_against_oracle(orc, engine, dup, "full", lanes=(0, "r1", "mono", "gen", 1, 4))


def test_early_out_e1_keeps_every_flag_and_every_contact(W, orc, engine):
    """GJKEPA_TUNE_EARLY_OUT = 1 (shortcut E1, opt-in): a support value s.d < -1e-6 ends the walk of a separated pair.
    Against the oracle, which has no shortcut: flags identical on every pair and in every kernel family; pairs that do
    not take the shortcut keep their exit path and iteration count; contacts keep depth and normal."""
    cases = [(W.make_pairs("box8", 200000, 1.0, seed=41), 8, "epa"), (W.make_pairs("hull", 60000, 0.85, seed=42, n_vert=32), 32, "epa"),
             (W.make_mixed(40000, 0.6, seed=43), 0, "epa"), (W.make_mixed(40000, 0.5, seed=44, n_lo=1, n_hi=5), 0, "gjk"),
             (W.make_shared_store(20000, 150000, n_vert=16, seed=45), 16, "full"), (W.make_touching_mix(40000, seed=46), 20, "epa"),
             (W.make_pairs("hull", 20000, 1.5, seed=47, n_vert=100), 0, "gjk")]
    for b, hint, mode in cases:
        ref = orc.batch(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=orc.MODE_EPA)
        for variant in (0, "mono", 1, 4):
            configure(engine, variant)
            engine.set_tuning(engine.TUNE_EARLY_OUT, 1)
            engine.set_max_hull_vertices(hint)
            got = engine.batch_host(b.verts, b.hull_off, b.pair_a, b.pair_b, mode=mode)
            assert np.array_equal(got["collision"], ref["collision"]), (b.name, variant)
            e1 = (got["status"] & 0xff) == 13
            sph = (ref["status"] & 0xff) == 1                     # E1 also drops the sphere test: those pairs walk on
            assert e1.sum() > 0 and not ref["collision"][e1].any()
            keep = ~e1 & ~sph
            assert np.array_equal(got["status"][keep] & 0xffff, ref["status"][keep] & 0xffff)
            if mode != "gjk":
                hit = (ref["collision"] > 0) & ((ref["status"] >> 24) == 0)
                assert np.all(np.abs(got["depth"][hit] - ref["depth"][hit]) <= DEPTH_RTOL * np.maximum(ref["depth"][hit], 1.0))
                assert np.all(got["depth"][~(ref["collision"] > 0)] == 0.0)
    engine.set_tuning(engine.TUNE_EARLY_OUT, 0)
    engine.set_max_hull_vertices(0)


def _minimising_facets(p1, p2, tol=1e-9):
    """Qhull hull of all pairwise differences (independent of the oracle): depth of the origin and the unit normals of
    every facet plane that attains it to within `tol` -- with tied depths the reference reports whichever of them its
    (un-vendored) QuickHull happens to list first, so any of them is a correct answer and nothing else is."""
    from scipy.spatial import ConvexHull
    md = (p1[:, None, :] - p2[None, :, :]).reshape(-1, 3)
    eq = ConvexHull(md).equations
    d = -eq[:, 3]
    d0 = float(d.min())
    return d0, eq[d <= d0 + tol, :3]


def test_tied_depth_normals_are_minimising_facets_of_the_minkowski_hull(W, engine):
    """Box-box face contacts -- the commonest real case -- have exactly tied depths (two or three faces of the Minkowski
    difference at the same distance); there the comparison with the oracle can only be by |cos| or skipped.  Ground truth
    instead: the returned depth is THE minimum over the Qhull facets and the returned normal is ONE OF the facets that
    attain it, to 1e-9, for every kernel family."""
    rng = np.random.default_rng(41)
    cube = np.array([[sx, sy, sz] for sz in (-.5, .5) for sy in (-.5, .5) for sx in (-.5, .5)])
    hulls, pa, pb = [], [], []
    offs = [(0.5, 0.5, 0.0), (0.5, 0.5, 0.5), (0.25, 0.25, 0.0), (0.75, 0.75, 0.1), (0.3, 0.0, 0.3), (0.5, 0.0, 0.0), (0.9, 0.9, 0.9)]
    for k in range(700):                      # axis-aligned cubes of equal or different size: exact ties
        s1, s2 = (1.0, 1.0) if k % 2 else (1.0, float(rng.choice([0.5, 1.5, 2.0])))
        o = np.array(offs[k % len(offs)]) * rng.choice([-1.0, 1.0], size=3)
        hulls += [cube * s1, cube * s2 + o * 0.5 * (s1 + s2)]
        pa.append(2 * k); pb.append(2 * k + 1)
    b8 = W.make_pairs("box8", 700, 0.6, seed=43)          # rotated boxes: near-ties, edge and face contacts
    base = len(hulls)
    for h in range(b8.n_hulls):
        hulls.append(b8.hull(h))
    pa += (b8.pair_a + base).tolist(); pb += (b8.pair_b + base).tolist()
    verts = np.concatenate([np.ascontiguousarray(h.T).reshape(-1) for h in hulls])
    off = np.concatenate([[0], np.cumsum([len(h) for h in hulls])]).astype(np.int64)
    pa, pb = np.array(pa, np.int32), np.array(pb, np.int32)
    for L in (0, "gen", 4):
        configure(engine, L)
        got = engine.batch_host(verts, off, pa, pb, mode="epa")
        hit = np.nonzero((got["collision"] > 0) & ((got["status"] >> 24) == 0))[0]
        assert len(hit) > 700
        for i in hit:
            d0, facets = _minimising_facets(hulls[pa[i]], hulls[pb[i]])
            assert abs(got["depth"][i] - max(d0, 0.0)) <= 1e-9 * max(d0, 1.0), (L, int(i), got["depth"][i], d0)
            cosd = facets @ got["normal"][i]
            best = np.max(np.abs(cosd)) if d0 < 1e-9 else np.max(cosd)      # touching: the sign follows the face winding
            assert 1.0 - best <= 1e-9, (L, int(i), got["normal"][i].tolist(), facets.tolist())
    configure(engine, 0)
