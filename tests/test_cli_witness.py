"""CPU: cli.calculate_witness restates circuitcompiler.Circuit.CalculateWitness (circuit.go:158-186);
it must reproduce the witnesses the Go binary printed (and circuit_test.go:81-82)."""
import json
import os


def test_calculate_witness_matches_go(golden_dir):
    from gosnark_b200.cli import calculate_witness
    for name in ("x3x5", "mul", "chain21"):
        g = json.load(open(os.path.join(golden_dir, f"gobin_{name}.json")))
        assert calculate_witness(g["compiledcircuit"], g["private"], g["public"]) == g["witness"]   # circuit_test.go:81-82


