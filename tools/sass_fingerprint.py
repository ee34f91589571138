"""Fingerprint of a kernel's SASS instruction stream (opcodes + operands, addresses and encodings stripped):
    python tools/sass_fingerprint.py sprc_b200/csrc/build/gemm2.o [substring of the mangled kernel name]
Used to show that adding the FOLD = 1 instantiation of the CTA-pair GEMM left the FOLD = 0 kernel (the one every
default path launches) instruction-for-instruction unchanged: profiles/r01r_gemm2_sass_fingerprint.txt."""
import hashlib
import re
import subprocess
import sys


def kernels(obj):
    out, cur = {}, None
    for line in subprocess.run(["cuobjdump", "-sass", obj], check=True, capture_output=True, text=True).stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m and cur:
            out[cur].append(m.group(1).strip())
    return out


if __name__ == "__main__":
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    for name, ins in kernels(sys.argv[1]).items():
        if want in name:
            ops = [re.sub(r"\bU?R\d+\b|\bU?P\d+\b", "r", i) for i in ins]   # register names masked
            print(f"{len(ins):6d} instructions  sha256 {hashlib.sha256(chr(10).join(ins).encode()).hexdigest()[:16]}"
                  f"  registers-masked sha256 {hashlib.sha256(chr(10).join(ops).encode()).hexdigest()[:16]}  {name}")
