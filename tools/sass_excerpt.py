"""SASS evidence for profiles/: per kernel of libsom_b200.so, how many tcgen05 / TMA / bulk-reduction instructions it holds.

    python tools/sass_excerpt.py > profiles/r2_sass_excerpt.md
"""
import collections
import re
import subprocess

SO = "xpysom_dask_b200/libsom_b200.so"
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKRED", "UTMAREDG", "UCGABAR", "SYNCS", "FMNMX3", "REDG", "ACQBULK"]
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
kern, counts, samples = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[kern] = collections.Counter()
        samples[kern] = {}
        continue
    if kern is None:
        continue
    for k in KEYS:
        if re.search(r"\b%s[\.\s]" % k, line):
            counts[kern][k] += 1
            ins = re.sub(r"/\*[0-9a-fx]+\*/", "", line).strip().rstrip(";").strip()
            samples[kern].setdefault(k, ins)
print("# SASS of the shipped `%s` (cuobjdump -sass, sm_100a): tensor-core / TMA / bulk-reduction instructions per kernel\n" % SO)
print("| kernel | " + " | ".join(KEYS) + " |\n|---|" + "---:|" * len(KEYS))
for k, c in counts.items():
    if sum(c.values()):
        print("| `%s` | %s |" % (k.replace("void ", ""), " | ".join(str(c[x]) if c[x] else "" for x in KEYS)))
print("\nFirst occurrence of each mnemonic in the fp16 kernels:\n\n```")
for k, smp in samples.items():
    if "bmu_tc3" in k and smp:
        print(k.replace("void ", ""))
        for key, ins in smp.items():
            print("    %s" % ins[:150])
print("```")
