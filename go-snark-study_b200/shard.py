"""Index-range sharding of a Groth16 proving key across ranks — the same
arithmetic as groth16_pk_load in csrc/prove_host.cuh: rank g of P owns
[m*g/P, m*(g+1)/P) of At / B1 / B2, the part of it above NPublic of BACDelta, and the
same fraction of PowersTauDelta; the blinding points ride on rank 0."""


def shard_ranges(m, npublic, n_ptd, rank, world):
    assert world >= 1 and 0 <= rank < world
    lo, hi = m * rank // world, m * (rank + 1) // world
    clo = min(max(lo, npublic + 1), hi)
    plo, phi = n_ptd * rank // world, n_ptd * (rank + 1) // world
    return {"lo": lo, "hi": hi, "clo": clo, "plo": plo, "phi": phi, "lead": rank == 0}
