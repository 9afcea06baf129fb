"""Which kernels changed?  Compares two `cuobjdump -sass libyttm_b200.so` dumps function by function (whitespace
normalised: cuobjdump aligns columns to the longest line of the whole file).  Used to prove that adding an
experimental kernel or a test-only #ifdef leaves the measured kernels' machine code untouched.
    cuobjdump -sass youtokentome_b200/libyttm_b200.so > /tmp/new.sass ; python tools/sass_diff.py /tmp/old.sass /tmp/new.sass"""
import re
import sys


def functions(path):
    parts = re.split(r"\n\s*Function : ", open(path).read())
    out = {}
    for p in parts[1:]:
        name, body = p.split("\n", 1)
        out[name.strip()] = "\n".join(re.sub(r"\s+", " ", ln).strip() for ln in body.splitlines())
    return out


def main():
    a, b = functions(sys.argv[1]), functions(sys.argv[2])
    for k in sorted(set(a) | set(b)):
        state = "only in new" if k not in a else "only in old" if k not in b else "same" if a[k] == b[k] else "CHANGED"
        print("%-12s %s" % (state, k[-90:]))
    return 0 if all(a[k] == b[k] for k in a if k in b) else 1


if __name__ == "__main__":
    sys.exit(main())
