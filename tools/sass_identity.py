"""Which kernels of the current build are instruction-for-instruction the kernels of an earlier commit?

    python tools/sass_identity.py fac4b2d > profiles/r01_sass_identity.txt

Builds csrc/ of the given commit in a scratch directory, dumps the SASS of both builds with cuobjdump and compares the
instruction streams kernel by kernel (addresses and encodings stripped).  Used at the end of round 1: everything
committed after the GPU budget was spent had to leave the kernels that HAVE run on hardware untouched; new code paths
are separate template instantiations / separate kernels behind switches that are off by default."""
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def funcs(path):
    text = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    out, cur = {}, None
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m2 = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(.*?)\s*/\* 0x[0-9a-f]+ \*/", line)
        if cur is not None and m2:
            out[cur].append(m2.group(1))
    return out


def demangle(names):
    p = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True)
    return dict(zip(names, p.stdout.splitlines()))


def key(demangled):
    # three kernels gained a defaulted bool template parameter after the reference commit; <..., false> IS the old kernel
    k = re.sub(r"^void\s+", "", demangled)
    k = re.sub(r"(igemm_nt_kernel<\d+, \d+), false>", r"\1>", k)
    k = re.sub(r"(wgrad_halo_kernel<\d+), false>", r"\1>", k)
    k = k.replace("igemm_nt_halo_kernel<false>", "igemm_nt_halo_kernel")
    return k


def collect(objdir):
    all_ = {}
    for f in glob.glob(os.path.join(objdir, "*.o")):
        all_.update(funcs(f))
    dm = demangle(list(all_))
    return {key(dm[n]): v for n, v in all_.items()}


def main():
    commit = sys.argv[1]
    with tempfile.TemporaryDirectory() as tmp:
        ar = subprocess.run(["git", "-C", ROOT, "archive", commit, "ecg-multimodal-model_b200/csrc", "include"],
                            capture_output=True, check=True).stdout
        subprocess.run(["tar", "-x", "-C", tmp], input=ar, check=True)
        subprocess.run(["make", "-C", os.path.join(tmp, "ecg-multimodal-model_b200", "csrc"), "-j", "16"],
                       capture_output=True, check=True)
        old = collect(os.path.join(tmp, "ecg-multimodal-model_b200", "csrc", "build"))
    new = collect(os.path.join(ROOT, "ecg-multimodal-model_b200", "csrc", "build"))
    same = sorted(k for k in old if k in new and old[k] == new[k])
    changed = sorted(k for k in old if k in new and old[k] != new[k])
    gone = sorted(k for k in old if k not in new)
    added = sorted(k for k in new if k not in old)
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    print(f"# kernels of {commit} against the build of {head}: {len(old)} kernels then, {len(same)} instruction-identical "
          f"now, {len(changed)} changed, {len(gone)} gone, {len(added)} new")
    for k in changed:
        print(f"CHANGED  {k}   ({len(old[k])} -> {len(new[k])} instructions)")
    for k in gone:
        print(f"GONE     {k}")
    for k in added:
        print(f"NEW      {k}   ({len(new[k])} instructions)")
    for k in same:
        print(f"same     {k}")


if __name__ == "__main__":
    main()
