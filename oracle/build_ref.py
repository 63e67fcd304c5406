"""Recipe for oracle/_ref: a verbatim, runnable copy of the reference's hot-path packages.

TEST / BASELINE INFRASTRUCTURE ONLY.  The reference (dogeplusplus/torch-actor-critic) is pure Python: there is
nothing to compile, "building" it means placing its own files where the GPU box can import them.  This script copies

    /root/reference/{buffer,networks,sac,environments}/*.py  ->  oracle/_ref/

byte for byte (never into git: oracle/_ref/ is git-ignored, but not gpurun-ignored, so it travels to the GPU box like a
built .so).  bench.py `--impl reference` then imports these files UNMODIFIED under the stubs of oracle/ref_stubs.py
and times `ReplayBuffer.sample -> SAC.update_critic -> SAC.update_policy -> update_targets`
(sac/algorithm.py:275-278) on the host cores (`cpu_baseline.kind = "reference"`).

    python oracle/build_ref.py            # run in the build container; __graft_entry__.build() calls it too
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
PACKAGES = ("buffer", "networks", "sac", "environments")


def available() -> bool:
    return os.path.isfile(os.path.join(OUT, "sac", "algorithm.py"))


def build(quiet: bool = False) -> bool:
    """Copy the reference packages.  Returns True when oracle/_ref is usable afterwards."""
    if not os.path.isdir(REF):
        return available()
    manifest = []
    for pkg in PACKAGES:
        src, dst = os.path.join(REF, pkg), os.path.join(OUT, pkg)
        os.makedirs(dst, exist_ok=True)
        for name in sorted(os.listdir(src)):
            if not name.endswith(".py"):
                continue
            shutil.copyfile(os.path.join(src, name), os.path.join(dst, name))
            with open(os.path.join(dst, name), "rb") as f:
                manifest.append(f"{hashlib.sha256(f.read()).hexdigest()}  {pkg}/{name}")
    with open(os.path.join(OUT, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    if not quiet:
        print(f"[build_ref] {len(manifest)} reference files -> {OUT}", file=sys.stderr)
    return available()


if __name__ == "__main__":
    ok = build()
    sys.exit(0 if ok else 1)
