"""GPU: the multi-GPU decomposition run on ONE device — two index-range shards
of the same proving key (rank 0/2 and 1/2), partial records concatenated as the
NCCL all-gather would, b200_groth16_finalize_device — must reproduce the
unsharded proof and the known-discrete-log expectation."""
import numpy as np
import pytest

from oracle import ref_py as o

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("logn,world,weights", [(6, 2, None), (8, 3, None), (10, 8, None), (10, 8, (85, 275)), (8, 3, (130, 330))])
def test_sharded_equals_single(logn, world, weights):
    """weights: the partition weights x100 (B200_CFG_SHARD_W_AB / _W_G2) — other cuts, same proof; every rank's cut is the one
    the Python mirror of csrc/shard_partition.h computes (b200_groth16_shard_info)."""
    import torch
    from gosnark_b200 import _lib, shard
    from gosnark_b200._lib import check, ints_to_limbs, lib, ptr
    from gosnark_b200.bn128 import _unflatten_g1, _unflatten_g2
    from gosnark_b200.synthetic import SyntheticGroth16
    _lib.init()
    L = lib()
    syn = SyntheticGroth16(logn)
    m, npx = syn.m, 2 * syn.n - 1
    r_l, s_l = ints_to_limbs([syn.r]), ints_to_limbs([syn.s])
    d_w = torch.from_numpy(syn.w.view(np.int64)).cuda()
    d_px = torch.from_numpy(syn.px.view(np.int64)).cuda()
    torch.cuda.synchronize()

    def read(d):
        out = d.cpu().numpy().view(np.uint64)
        pa, pc = _unflatten_g1(out[:24])
        return pa, _unflatten_g2(out[24:])[0], pc

    pk1 = syn.load_pk(0, 1)
    d_single = torch.zeros(48, dtype=torch.int64, device="cuda")
    check(L.b200_groth16_prove_device(pk1, d_w.data_ptr(), m, d_px.data_ptr(), npx, ptr(r_l), ptr(s_l),
                                      d_single.data_ptr(), None))
    parts = torch.zeros(128 * world, dtype=torch.int64, device="cuda")
    w_ab, w_g2 = weights or (100, 280)
    check(L.b200_config(_lib.CFG_SHARD_W_AB, w_ab))
    check(L.b200_config(_lib.CFG_SHARD_W_G2, w_g2))
    shard.W_AB, shard.W_G2 = w_ab / 100.0, w_g2 / 100.0
    try:
        pks = [syn.load_pk(rk, world) for rk in range(world)]
        for rk, pk in enumerate(pks):      # the library cut the key exactly as the Python mirror says
            info = np.zeros(12, dtype=np.uint64)
            check(L.b200_groth16_shard_info(pk, ptr(info)))
            sets = shard.shard_ranges(m, syn.npublic, syn.n_ptd, rk, world)["sets"]
            assert [int(x) for x in info[:6]] == [v for k in range(3) for v in (sets[k]["lo"], sets[k]["hi"])]
    finally:
        check(L.b200_config(_lib.CFG_SHARD_W_AB, 100))
        check(L.b200_config(_lib.CFG_SHARD_W_G2, 280))
        shard.W_AB, shard.W_G2 = 1.0, 2.8
    for rk, pk in enumerate(pks):
        check(L.b200_groth16_prove_device(pk, d_w.data_ptr(), m, d_px.data_ptr(), npx, ptr(r_l), ptr(s_l),
                                          parts[128 * rk:].data_ptr(), None))
    d_out = torch.zeros(48, dtype=torch.int64, device="cuda")
    check(L.b200_groth16_finalize_device(pks[0], parts.data_ptr(), world, ptr(r_l), ptr(s_l), d_out.data_ptr(), None))
    check(L.b200_profile_read((__import__("ctypes").c_double * 8)()))      # device sync
    G1, G2 = o.BN.G1, o.BN.G2
    a1, b1, c1 = read(d_single)
    a2, b2, c2 = read(d_out)
    assert G1.affine(a1) == G1.affine(a2) and G2.affine(b1) == G2.affine(b2) and G1.affine(c1) == G1.affine(c2)
    ea, eb, ec = syn.expected_dlogs()
    assert G1.affine(a2) == G1.affine(G1.mul_scalar(G1.G, ea))
    assert G2.affine(b2) == G2.affine(G2.mul_scalar(G2.G, eb))
    assert G1.affine(c2) == G1.affine(G1.mul_scalar(G1.G, ec))
    # host-pointer API refuses a sharded key
    with pytest.raises(_lib.B200Error):
        pa, pb, pc = (np.zeros(k, dtype=np.uint64) for k in (12, 24, 12))
        check(L.b200_groth16_prove(pks[1], ptr(syn.w), m, ptr(syn.px), npx, ptr(r_l), ptr(s_l), ptr(pa), ptr(pb), ptr(pc)))
    for pk in pks + [pk1]:
        check(L.b200_pk_free(pk))


def test_synthetic_matches_oracle_prove():
    """The synthetic key as a groth16.Pk dict through the reference-order oracle gives the same proof
    (ties the known-dlog expectation to groth16.go:225-278 itself), n = 8."""
    from gosnark_b200 import _lib, groth16
    from gosnark_b200._lib import limbs_to_ints
    from gosnark_b200.synthetic import SyntheticGroth16
    _lib.init()
    syn = SyntheticGroth16(3)
    pk = syn.pk_dict()
    w, px = limbs_to_ints(syn.w), limbs_to_ints(syn.px)
    ref, raw = o.groth16_prove(syn.m, syn.npublic, pk, w, px, syn.r, syn.s)
    assert raw["hx"] == limbs_to_ints(syn.h0)
    ours = groth16.GenerateProofs({"NVars": syn.m, "NPublic": syn.npublic}, pk, w, px, r=syn.r, s=syn.s)
    G1, G2 = o.BN.G1, o.BN.G2
    for k, G in (("PiA", G1), ("PiB", G2), ("PiC", G1)):
        assert G.affine(ours[k]) == G.affine(ref[k])
    ea, eb, ec = syn.expected_dlogs()
    assert G1.affine(ref["PiA"]) == G1.affine(G1.mul_scalar(G1.G, ea))
    assert G1.affine(ref["PiC"]) == G1.affine(G1.mul_scalar(G1.G, ec))


@pytest.mark.parametrize("world", [2, 3, 8])
def test_pinocchio_sharded_equals_go_binary(golden_dir, world):
    """Sharded Pinocchio keys (whole MSMs dealt to the ranks, b200_pinocchio_pk_load_shard): every rank's 2 KB record,
    concatenated as the NCCL all-gather would and summed by b200_pinocchio_finalize_records, reproduces the Go binary's
    proof (snark.go:254-289) on affine coordinates; the host-pointer call refuses a sharded key without a communicator."""
    import json
    import os
    import torch
    from gosnark_b200 import _lib
    from gosnark_b200._lib import check, ints_to_limbs, lib, ptr
    from gosnark_b200.bn128 import R, _flatten_g1, _flatten_g2, _unflatten_g1, _unflatten_g2, reduce_scalar
    def t3(p):
        return tuple(p)

    def pinocchio_pk(setup):
        pk = {k: [t3(p) for p in setup["Pk"][k]] for k in ("A", "C", "Kp", "Ap", "Bp", "Cp")}
        pk["B"] = [tuple(tuple(c) for c in p) for p in setup["Pk"]["B"]]
        pk["Z"] = setup["Pk"]["Z"]
        pk["G1T"] = [t3(p) for p in setup["G1T"]]
        return pk
    _lib.init()
    L = lib()
    with open(os.path.join(golden_dir, "gobin_chain21.json")) as f:
        g = json.load(f)
    cc = g["compiledcircuit"]
    pk = pinocchio_pk(g["pinocchio_setup"])
    m, npub = cc["NVars"], cc["NPublic"]
    arr = {k: _flatten_g1(pk[k][:m]) for k in ("A", "Ap", "Bp", "C", "Cp", "Kp")}
    b2, g1t = _flatten_g2(pk["B"][:m]), _flatten_g1(pk["G1T"])
    z = ints_to_limbs([int(x) % R for x in pk["Z"]])
    w = ints_to_limbs([reduce_scalar(x) for x in g["witness"]])
    px = ints_to_limbs([int(x) % R for x in g["px"]])
    recs = torch.zeros(256 * world, dtype=torch.int64, device="cuda")          # world x 2 KB
    handles = []
    for rk in range(world):
        h = _lib._h(0)
        check(L.b200_pinocchio_pk_load_shard(ptr(arr["A"]), ptr(arr["Ap"]), ptr(b2), ptr(arr["Bp"]), ptr(arr["C"]), ptr(arr["Cp"]),
                                             ptr(arr["Kp"]), m, ptr(g1t), len(pk["G1T"]), ptr(z), len(pk["Z"]), npub, 0, rk, world, h))
        handles.append(h.value)
        check(L.b200_pinocchio_prove_record_device(h.value, ptr(w), m, ptr(px), px.shape[0], recs[256 * rk:].data_ptr()))
    out_g1, out_b = np.zeros(84, dtype=np.uint64), np.zeros(24, dtype=np.uint64)
    check(L.b200_pinocchio_finalize_records(recs.data_ptr(), world, ptr(out_g1), ptr(out_b)))
    proof = dict(zip(("PiA", "PiAp", "PiBp", "PiC", "PiCp", "PiH", "PiKp"), _unflatten_g1(out_g1)))
    proof["PiB"] = _unflatten_g2(out_b)[0]
    ref = g["pinocchio_proofs"]
    for k in ("PiA", "PiAp", "PiBp", "PiC", "PiCp", "PiH", "PiKp"):
        assert o.BN.G1.affine(proof[k]) == o.BN.G1.affine(tuple(ref[k])), k
    assert o.BN.G2.affine(proof["PiB"]) == o.BN.G2.affine(tuple(tuple(c) for c in ref["PiB"]))
    with pytest.raises(_lib.B200Error):          # no communicator: the host-pointer call cannot finish a sharded proof
        check(L.b200_pinocchio_prove(handles[0], ptr(w), m, ptr(px), px.shape[0], ptr(out_g1), ptr(out_b)))
    for h in handles:
        check(L.b200_pk_free(h))
