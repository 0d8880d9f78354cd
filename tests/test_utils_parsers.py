"""CPU: the string / hex data formats of the reference's utils package, pinned on the wasm demo's
fixtures (wasm/index.js:2-8, committed as tests/golden/wasm_index_strings.json by oracle/make_wasm_golden.py): the
string-form circuit, Pinocchio setup and px parse to the integers that give the K5 proof of SURVEY §8c."""
import json
import os
import random

import pytest

import gosnark_b200  # noqa: F401
from gosnark_b200 import utils
from oracle import ref_py as o

G1, G2 = o.BN.G1, o.BN.G2

# the Go binary's proof for this fixture (affine coordinates)
K5 = {
    "PiA": (4453507680665149551040865562645380579364190471901699692881203134836874519542,
            18327733998491295152058637956806021119992767714079462162415761243874009797647),
    "PiAp": (19271241879650057039755054804130298949564247997614523192956793508359182605457,
             9682911175591685512225940843004636890131169135056577733106588795772383802682),
    "PiBp": (12630797189265093320706291914826823023280647637620423582842960093100457835952,
             5933239191911630044333497325970788294993348802214774414718229652177273836113),
    "PiC": (7932023228249558612977583339058349538139042298887658268658886774970191843296,
            13410307163435450492910109421154888355324119085192555778875795082615307140659),
    "PiCp": (18290031543457093605365496304949211929513190764884777571864773072014926354289,
             18129879325255437258799480545885342266167716944513601035559429974422100868703),
    "PiH": (5526923620061579272744927845013801971835074326128480358904962217694798146354,
            13129959477742735457882063358869968204804042081618513582088963896408716291012),
    "PiKp": (9754005701745928147331140462545295587820355778153788654448357986546636384570,
             7629956750043724324672049348589299295850750268696876445917488213146611981806),
}
K5_PIB = ((12747450872952006421862372390439318369521376181718515375338448395011170850826,
           20842821080997740635712736278207361145752931843939505354148285977982621601971),
          (5513640086630407359706325037824666136883912832971027822836090446505241620746,
           282422234443388129371664703631910913162297635394986724152704431892082583326))


@pytest.fixture(scope="module")
def wasm(golden_dir):
    return json.load(open(os.path.join(golden_dir, "wasm_index_strings.json")))


def test_wasm_fixture_parses_to_the_k5_proof(wasm):
    circuit = utils.CircuitFromString(wasm["circuit"])
    setup = utils.SetupFromString(wasm["setup"])
    px = utils.ArrayStringToBigInt(wasm["px"])
    assert circuit["NVars"] == 8 and circuit["R1CS"]["A"][3][0] == 5 and circuit["Witness"] == []
    assert setup["Pk"]["G1T"][0] == (1, 2, 1) and len(setup["Pk"]["Z"]) == 7
    w = [1, 35, 3, 9, 27, 30, 35, 1]
    proof, _ = o.pinocchio_prove(8, 1, setup["Pk"], w, px)
    for k, v in K5.items():
        assert G1.affine(proof[k])[:2] == v, k
    assert G2.affine(proof["PiB"])[:2] == K5_PIB
    # and back: the string form of what was parsed is the fixture (modulo the top-level G1T of the older layout)
    back = utils.SetupToString(setup)
    assert back["Pk"]["A"] == wasm["setup"]["Pk"]["A"] and back["Pk"]["B"] == wasm["setup"]["Pk"]["B"]
    assert back["Vk"] == wasm["setup"]["Vk"] and back["Pk"]["G1T"] == wasm["setup"]["G1T"]
    assert utils.ArrayBigIntToString(px) == wasm["px"]
    assert utils.CircuitToString(circuit)["R1CS"] == wasm["circuit"]["R1CS"]


def test_round_trips_all_forms():
    rng = random.Random(3)
    fe = lambda: rng.randrange(o.Q)
    p1 = lambda: (fe(), fe(), fe())
    p2 = lambda: ((fe(), fe()), (fe(), fe()), (fe(), fe()))
    sproof = {k: (p2() if k == "PiB" else p1()) for k in ("PiA", "PiAp", "PiB", "PiBp", "PiC", "PiCp", "PiH", "PiKp")}
    gproof = {"PiA": p1(), "PiB": p2(), "PiC": p1()}
    gsetup = {"Pk": {"BACDelta": [(0, 0, 0), p1()], "Z": [fe(), 1], "PowersTauDelta": [p1()],
                     "G1": {"Alpha": p1(), "Beta": p1(), "Delta": p1(), "At": [p1(), p1()], "BACGamma": [p1()]},
                     "G2": {"Beta": p2(), "Gamma": p2(), "Delta": p2(), "BACGamma": [p2(), p2()]}},
              "Vk": {"IC": [p1(), p1()], "G1": {"Alpha": p1()}, "G2": {"Beta": p2(), "Gamma": p2(), "Delta": p2()}}}
    assert utils.ProofFromString(utils.ProofToString(sproof)) == sproof
    assert utils.ProofFromHex(utils.ProofToHex(sproof)) == sproof
    assert utils.GrothProofFromString(utils.GrothProofToString(gproof)) == gproof
    assert utils.GrothProofFromHex(utils.GrothProofToHex(gproof)) == gproof
    assert utils.GrothSetupFromString(utils.GrothSetupToString(gsetup)) == gsetup
    assert utils.GrothSetupFromHex(utils.GrothSetupToHex(gsetup)) == gsetup
    assert utils.GrothVkFromString(utils.GrothSetupToString(gsetup)["Vk"]) == gsetup["Vk"]
    s = utils.GrothProofToHex(gproof)
    assert s["PiA"][0] == format(gproof["PiA"][0], "x") and not s["PiA"][0].startswith("0x")
    circ = {"NVars": 3, "NPublic": 1, "NSignals": 3, "PrivateInputs": ["a"], "PublicInputs": ["b"], "Signals": ["one", "b", "a"],
            "Constraints": [], "Witness": [1, 5, o.R - 1], "R1CS": {"A": [[0, -1, 2]], "B": [[1, 0, 0]], "C": [[0, 1, 0]]}}
    assert utils.CircuitFromString(utils.CircuitToString(circ)) == circ
    assert utils.CircuitFromHex(utils.CircuitToHex(circ)) == circ
    assert utils.CircuitToString(circ)["R1CS"]["A"] == [["0", "-1", "2"]]           # big.Int.String() keeps the sign
    assert utils.ArrayBigIntToHex([255, -255, 0]) == ["ff", "-ff", "0"]


@pytest.mark.parametrize("bad", ["", "12a", "0x10", " 5", "1_000", "١٢"])
def test_malformed_digit_strings_raise_like_setstring(bad):
    with pytest.raises(ValueError, match="error parsing px from pxString"):
        utils.ArrayStringToBigInt(["1", bad])
    with pytest.raises(ValueError):
        utils.String3ToBigInt(["1", bad, "1"])
    if bad != "12a":
        with pytest.raises(ValueError, match="error parsing px from pxHex"):
            utils.ArrayHexToBigInt([bad])
    with pytest.raises(ValueError):
        utils.String3ToBigInt(["1", "2"])
