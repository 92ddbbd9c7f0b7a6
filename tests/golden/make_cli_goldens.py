"""Generates tests/golden/cli_goldens.json: stdout / stderr / exit code of the UNMODIFIED reference CLI
(/root/reference/mainDriver.cu built with the reference's own flags, CPU path `-c`) on input files written from the
committed fixtures.  tests/test_dropin.py runs the same command lines with `-g` through tests/cpp/dropin_main.cpp linked
against libsa_b200.so and compares byte for byte.  Run in the build container (needs /root/reference and nvcc):

    python tests/golden/make_cli_goldens.py
"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cli_cases  # noqa: E402

REF = "/root/reference"


def main():
    exe = os.path.join(ROOT, "oracle", "_ref", "alignSequence_ref")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.run(["nvcc", "-std=c++14", "-m64", "--expt-relaxed-constexpr", "-include", "cstdint", "-w",
                    os.path.join(REF, "mainDriver.cu"), "-o", exe], check=True)
    out = []
    with tempfile.TemporaryDirectory() as tmp:
        cli_cases.write_inputs(tmp)
        for case in cli_cases.CASES:
            args = [a.replace("{dev}", "-c") for a in case["args"]]
            r = subprocess.run([exe] + args, cwd=tmp, capture_output=True)
            out.append(dict(name=case["name"], args=case["args"], returncode=r.returncode,
                            stdout=r.stdout.decode("latin1"), stderr=r.stderr.decode("latin1")))
            print(case["name"], r.returncode, len(r.stdout))
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "cli_goldens.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
