#!/usr/bin/env python
"""Static resource usage of every kernel in liblirec_b200.so (cuobjdump -res-usage; no GPU needed), plus the
SASS mnemonics that show the Blackwell-native paths: registers, spills (LOCAL / STACK), static shared memory.

    python tools/res_usage.py > profiles/rNN_resource_usage.txt
"""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(HERE, "lirec_b200", "liblirec_b200.so")
out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
names = re.findall(r"Function (\S+):\n\s*(REG:\d+ STACK:\d+ SHARED:\d+ LOCAL:\d+)", out)
dem = subprocess.run(["c++filt"], input="\n".join(n for n, _ in names), capture_output=True, text=True).stdout.split("\n")
print("# cuobjdump -res-usage lirec_b200/liblirec_b200.so  (sm_100a; SHARED = static only, the GEMM kernels take"
      " their ring buffers as dynamic shared memory)")
for d, (_, res) in zip(dem, names):
    print("%-46s %s" % (res, re.sub(r"\(.*", "", d)[:110]))
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
print("# SASS mnemonic counts")
for m in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", "LDGMC", "HMMA."):
    print("%-10s %d" % (m, len(re.findall(r"\b" + re.escape(m), sass.replace("UTCHMMA", "UTCHMMA_") if m == "HMMA." else sass))))
