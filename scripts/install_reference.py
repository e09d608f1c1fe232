#!/usr/bin/env python3
"""Install the UNMODIFIED reference (aidanmrli/rwm-pt-pytorch) into baseline/_ref -- the one offline install the task
allows -- so that it travels to the GPU box with the snapshot (baseline/_ref is git-ignored, not gpurun-ignored):

    python scripts/install_reference.py            # needs /root/reference (authoring container only)

  * the four packages (algorithms, interfaces, proposal_distributions, target_distributions) are pip-installed from a copy of
    the source tree (the reference tree is read-only and setuptools writes build files next to pyproject.toml);
  * the reference's own test scripts are copied verbatim to baseline/_ref/_reference_tests/ (its pyproject excludes
    tests from the wheel); tests/test_reference_suite.py runs them against the drop-in classes;
  * data/*.py (average_seeds.py, ...), the consumers of the sweep drivers' JSON schema, go to
    baseline/_ref/_reference_data_tools/.
Nothing under baseline/_ref is tracked by git or imported by the product."""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("RWMPT_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def main():
    if not os.path.isdir(os.path.join(REF, "algorithms")):
        print(f"reference not found at {REF}: nothing to do (baseline/_ref is used as it is)")
        return 0
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "ref")
        os.makedirs(src)
        for name in ("algorithms", "interfaces", "proposal_distributions", "target_distributions"):
            shutil.copytree(os.path.join(REF, name), os.path.join(src, name))
        for name in ("pyproject.toml", "README.md", "LICENSE"):
            if os.path.exists(os.path.join(REF, name)):
                shutil.copy(os.path.join(REF, name), src)
        if os.path.isdir(DST):
            shutil.rmtree(DST)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", DST, src]
        subprocess.run(cmd, check=True)
    tests_dst = os.path.join(DST, "_reference_tests")
    os.makedirs(tests_dst, exist_ok=True)
    for f in sorted(os.listdir(os.path.join(REF, "tests"))):
        if f.endswith(".py"):
            shutil.copy(os.path.join(REF, "tests", f), tests_dst)
    tools_dst = os.path.join(DST, "_reference_data_tools")       # data/average_seeds.py etc.: the consumers of the JSON schema
    os.makedirs(tools_dst, exist_ok=True)
    for f in sorted(os.listdir(os.path.join(REF, "data"))):
        if f.endswith(".py"):
            shutil.copy(os.path.join(REF, "data", f), tools_dst)
    print(f"installed the reference into {DST} (+ {len(os.listdir(tests_dst))} test scripts, {len(os.listdir(tools_dst))} data tools)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
