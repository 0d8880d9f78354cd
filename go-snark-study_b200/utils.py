"""Host mirror of the reference's ``utils`` package: the decimal-string and hex-string forms of scalars, points,
circuits, setups and proofs that its wasm wrapper and JS tooling exchange (utils/base10parsers.go, utils/hexparsers.go;
SURVEY §8f row 4).  Pure host code: these are the data formats either side of the prove path.

Shapes follow the Go types: ``[3]*big.Int`` = 3-tuple of ints (G1 Jacobian), ``[3][2]*big.Int`` = 3-tuple of 2-tuples
(G2), slices = lists.  A malformed digit string raises ValueError with the reference's error text; like
``big.Int.SetString(s, base)`` an optional sign is accepted, prefixes (0x), underscores and blanks are not."""
import re

_DEC = re.compile(r"^[+-]?[0-9]+$")
_HEX = re.compile(r"^[+-]?[0-9a-fA-F]+$")


def _parse(s, base, err):
    if not isinstance(s, str) or not (_DEC if base == 10 else _HEX).match(s):
        raise ValueError(err)
    return int(s, base)


def _dec(x):
    return str(int(x))                                   # big.Int.String()


def _hex(x):
    x = int(x)
    return ("-" if x < 0 else "") + format(abs(x), "x")   # fmt.Sprintf("%x", b)


# ---- []*big.Int -------------------------------------------------------------------------------- base10parsers.go:13-30
def ArrayBigIntToString(b):
    return [_dec(x) for x in b]


def ArrayStringToBigInt(s):
    return [_parse(x, 10, "error parsing px from pxString") for x in s]


def ArrayBigIntToHex(b):                                 # hexparsers.go:14-31
    return [_hex(x) for x in b]


def ArrayHexToBigInt(s):
    return [_parse(x, 16, "error parsing px from pxHex") for x in s]


# ---- [3]*big.Int ------------------------------------------------------------------------------- base10parsers.go:33-70
def String3ToBigInt(s):
    _len(s, 3)
    return tuple(_parse(x, 10, "error parsing [3]*big.Int from [3]string") for x in s)


def BigInt3ToString(b):
    _len(b, 3)
    return [_dec(x) for x in b]


def Array3StringToBigInt(s):
    return [String3ToBigInt(x) for x in s]


def Array3BigIntToString(b):
    return [BigInt3ToString(x) for x in b]


def Hex3ToBigInt(s):                                     # hexparsers.go:34-71
    _len(s, 3)
    return tuple(_parse(x, 16, "error parsing [3]*big.Int from [3]string") for x in s)


def BigInt3ToHex(b):
    _len(b, 3)
    return [_hex(x) for x in b]


def Array3HexToBigInt(s):
    return [Hex3ToBigInt(x) for x in s]


def Array3BigIntToHex(b):
    return [BigInt3ToHex(x) for x in b]


# ---- [2] and [3][2]*big.Int -------------------------------------------------------------------- base10parsers.go:72-133
def String2ToBigInt(s):
    _len(s, 2)
    return tuple(_parse(x, 10, "error parsing [2]*big.Int from [2]string") for x in s)


def String32ToBigInt(s):
    _len(s, 3)
    return tuple(String2ToBigInt(x) for x in s)


def BigInt32ToString(b):
    _len(b, 3)
    return [[_dec(c) for c in _len(x, 2)] for x in b]


def Array32StringToBigInt(s):
    return [String32ToBigInt(x) for x in s]


def Array32BigIntToString(b):
    return [BigInt32ToString(x) for x in b]


def Hex2ToBigInt(s):                                     # hexparsers.go:73-134
    _len(s, 2)
    return tuple(_parse(x, 16, "error parsing [2]*big.Int from [2]string") for x in s)


def Hex32ToBigInt(s):
    _len(s, 3)
    return tuple(Hex2ToBigInt(x) for x in s)


def BigInt32ToHex(b):
    _len(b, 3)
    return [[_hex(c) for c in _len(x, 2)] for x in b]


def Array32HexToBigInt(s):
    return [Hex32ToBigInt(x) for x in s]


def Array32BigIntToHex(b):
    return [BigInt32ToHex(x) for x in b]


# ---- [][]*big.Int ------------------------------------------------------------------------------ base10parsers.go:275-291
def ArrayArrayBigIntToString(b):
    return [ArrayBigIntToString(x) for x in b]


def ArrayArrayStringToBigInt(s):
    return [ArrayStringToBigInt(x) for x in s]


def ArrayArrayBigIntToHex(b):                            # hexparsers.go:277-293
    return [ArrayBigIntToHex(x) for x in b]


def ArrayArrayHexToBigInt(s):
    return [ArrayHexToBigInt(x) for x in s]


def _len(v, n):
    if len(v) != n:
        raise ValueError(f"expected {n} components, got {len(v)}")
    return v


# ---- structs: a (field -> kind) schema per Go struct, applied in either direction ---------------------------------
_S, _A, _P1, _A1, _P2, _A2, _AA = "scalar", "[]", "[3]", "[][3]", "[3][2]", "[][3][2]", "[][]"
_TO = {10: {_A: ArrayBigIntToString, _P1: BigInt3ToString, _A1: Array3BigIntToString, _P2: BigInt32ToString,
            _A2: Array32BigIntToString, _AA: ArrayArrayBigIntToString},
       16: {_A: ArrayBigIntToHex, _P1: BigInt3ToHex, _A1: Array3BigIntToHex, _P2: BigInt32ToHex,
            _A2: Array32BigIntToHex, _AA: ArrayArrayBigIntToHex}}
_FROM = {10: {_A: ArrayStringToBigInt, _P1: String3ToBigInt, _A1: Array3StringToBigInt, _P2: String32ToBigInt,
              _A2: Array32StringToBigInt, _AA: ArrayArrayStringToBigInt},
         16: {_A: ArrayHexToBigInt, _P1: Hex3ToBigInt, _A1: Array3HexToBigInt, _P2: Hex32ToBigInt,
              _A2: Array32HexToBigInt, _AA: ArrayArrayHexToBigInt}}

_SNARK_PK = {"G1T": _A1, "A": _A1, "B": _A2, "C": _A1, "Kp": _A1, "Ap": _A1, "Bp": _A1, "Cp": _A1, "Z": _A}   # snark.go:16-26
_SNARK_VK = {"Vka": _P2, "Vkb": _P1, "Vkc": _P2, "IC": _A1, "G1Kbg": _P1, "G2Kbg": _P2, "G2Kg": _P2, "Vkz": _P2}
_SNARK_PROOF = {"PiA": _P1, "PiAp": _P1, "PiB": _P2, "PiBp": _P1, "PiC": _P1, "PiCp": _P1, "PiH": _P1, "PiKp": _P1}
_GROTH_PK = {"BACDelta": _A1, "Z": _A, "PowersTauDelta": _A1,                                            # groth16.go:15-32
             "G1": {"Alpha": _P1, "Beta": _P1, "Delta": _P1, "At": _A1, "BACGamma": _A1},
             "G2": {"Beta": _P2, "Gamma": _P2, "Delta": _P2, "BACGamma": _A2}}
_GROTH_VK = {"IC": _A1, "G1": {"Alpha": _P1}, "G2": {"Beta": _P2, "Gamma": _P2, "Delta": _P2}}
_GROTH_PROOF = {"PiA": _P1, "PiB": _P2, "PiC": _P1}


def _apply(schema, obj, table):
    out = {}
    for k, kind in schema.items():
        if isinstance(kind, dict):
            out[k] = _apply(kind, obj[k], table)
        else:
            out[k] = table[kind](obj[k])
    return out


def _setup_to(setup, pk_schema, vk_schema, base):
    """SetupToString / GrothSetupToString: public parts only — the Toxic values are not carried (base10parsers.go:160-179)."""
    return {"Pk": _apply(pk_schema, setup["Pk"], _TO[base]), "Vk": _apply(vk_schema, setup["Vk"], _TO[base])}


def _setup_from(s, pk_schema, vk_schema, base):
    pk = dict(s["Pk"])
    if "G1T" in pk_schema and "G1T" not in pk and "G1T" in s:      # the wasm demo's older layout keeps G1T at top level
        pk["G1T"] = s["G1T"]
    return {"Pk": _apply(pk_schema, pk, _FROM[base]), "Vk": _apply(vk_schema, s["Vk"], _FROM[base])}


def SetupToString(setup):                                # base10parsers.go:160-179
    return _setup_to(setup, _SNARK_PK, _SNARK_VK, 10)


def SetupFromString(s):                                  # :181-273
    return _setup_from(s, _SNARK_PK, _SNARK_VK, 10)


def SetupToHex(setup):                                   # hexparsers.go:162-181
    return _setup_to(setup, _SNARK_PK, _SNARK_VK, 16)


def SetupFromHex(s):                                     # :183-275
    return _setup_from(s, _SNARK_PK, _SNARK_VK, 16)


def GrothSetupToString(setup):                           # base10parsers.go:435-454
    return _setup_to(setup, _GROTH_PK, _GROTH_VK, 10)


def GrothSetupFromString(s):                             # :481-559
    return _setup_from(s, _GROTH_PK, _GROTH_VK, 10)


def GrothVkFromString(s):                                # :456-479
    return _apply(_GROTH_VK, s, _FROM[10])


def GrothSetupToHex(setup):                              # hexparsers.go:437-456
    return _setup_to(setup, _GROTH_PK, _GROTH_VK, 16)


def GrothSetupFromHex(s):                                # :458-536
    return _setup_from(s, _GROTH_PK, _GROTH_VK, 16)


def ProofToString(p):                                    # base10parsers.go:349-359
    return _apply(_SNARK_PROOF, p, _TO[10])


def ProofFromString(s):                                  # :361-398
    return _apply(_SNARK_PROOF, s, _FROM[10])


def ProofToHex(p):                                       # hexparsers.go:351-361
    return _apply(_SNARK_PROOF, p, _TO[16])


def ProofFromHex(s):                                     # :363-400
    return _apply(_SNARK_PROOF, s, _FROM[16])


def GrothProofToString(p):                               # base10parsers.go:561-566
    return _apply(_GROTH_PROOF, p, _TO[10])


def GrothProofFromString(s):                             # :568-585
    return _apply(_GROTH_PROOF, s, _FROM[10])


def GrothProofToHex(p):                                  # hexparsers.go:538-543
    return _apply(_GROTH_PROOF, p, _TO[16])


def GrothProofFromHex(s):                                # :545-562
    return _apply(_GROTH_PROOF, s, _FROM[16])


_CIRCUIT_COPY = ("NVars", "NPublic", "NSignals", "PrivateInputs", "PublicInputs", "Signals", "Constraints")


def _circuit(c, table, witness_kind):
    out = {k: c.get(k) for k in _CIRCUIT_COPY}
    w = c.get("Witness")
    out["Witness"] = table[_A](w) if w else ([] if witness_kind == "from" else None)
    out["R1CS"] = {m: table[_AA](c["R1CS"][m]) for m in ("A", "B", "C")}
    return out


def CircuitToString(c):                                  # base10parsers.go:293-306
    return _circuit(c, _TO[10], "to")


def CircuitFromString(cs):                               # :308-346
    return _circuit(cs, _FROM[10], "from")


def CircuitToHex(c):                                     # hexparsers.go:295-308
    return _circuit(c, _TO[16], "to")


def CircuitFromHex(cs):                                  # :310-348
    return _circuit(cs, _FROM[16], "from")
