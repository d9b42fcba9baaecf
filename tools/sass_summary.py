"""SASS evidence for the hot kernels of libfwi_b200.so (runs without a GPU): per kernel, registers / spills / shared memory
from `cuobjdump -res-usage` and instruction counts from `cuobjdump -sass` (TMA loads, mbarrier syncs, FMAs, vector
shared/global accesses).   python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "full_waveform_inversion_b200", "libfwi_b200.so")
HOT = ("fd2d_step_kernel", "fd2d_tb2_kernel", "fd3d_step_kernel", "mc_umma_kernel", "mc_umma_pack", "mc_eval_kernel", "mc_sample_kernel", "fd3d_slab_sync")
PAT = collections.OrderedDict([
    ("UTMALDG", r"\bUTMALDG"), ("UTCHMMA(tcgen05.mma)", r"\bUTCHMMA"), ("LDTM(tcgen05.ld)", r"\bLDTM"), ("UTCBAR(tcgen05.commit)", r"\bUTCBAR"),
    ("UTCATOMSWS(tmem alloc)", r"\bUTCATOMSWS"), ("FMNMX3", r"\bFMNMX3"), ("SYNCS(mbarrier)", r"\bSYNCS"), ("FFMA", r"\bFFMA"), ("FADD", r"\bFADD"), ("FMUL", r"\bFMUL"),
    ("LDS.128", r"\bLDS\.(U\.)?128"), ("LDS", r"\bLDS"), ("LDG.128", r"\bLDG\.E\.(\w+\.)*128"), ("LDG", r"\bLDG"),
    ("STG.128", r"\bSTG\.E\.(\w+\.)*128"), ("STG", r"\bSTG"), ("ATOM/RED", r"\b(ATOMG|RED|ATOMS)\b"), ("BAR", r"\bBAR\."),
    ("ACQBULK", r"ACQBULK"), ("DFMA", r"\bDFMA"), ("MUFU", r"\bMUFU")])
res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
usage = {}
for m in re.finditer(r"Function (\S+):\n\s+REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res):
    usage[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)), int(m.group(5)))
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", sass)[1:]
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
print("SASS summary of %s (sm_100a); %d kernels in the library" % (os.path.relpath(LIB, ROOT), len(usage)))
tot = collections.Counter()
for b in blocks:
    name = b.split("\n", 1)[0].strip()
    body = b.split("\n", 1)[1] if "\n" in b else ""
    for k, p in PAT.items():
        tot[k] += len(re.findall(p, body))
    if not any(h in name for h in HOT):
        continue
    if "mc_eval_kernel" in name and "mc_eval_kernelILi9ELi1ELi8E" not in name:
        continue                                  # ~90 instantiations: list the default shape (C = 9, one medium, 8 samples per lane)
    reg, stack, shared, local = usage.get(name, (0, 0, 0, 0))
    counts = ["%s %d" % (k, len(re.findall(p, body))) for k, p in PAT.items() if re.search(p, body)]
    ninstr = len(re.findall(r"^\s+/\*[0-9a-f]{4}\*/", body, re.M))
    d = demangle(name)
    d = re.sub(r"\(CUtensorMap_st.*", "", d).replace("fwi::", "")
    print("\n%s\n  registers %d  stack %d B  static smem %d B  local %d B  instructions %d\n  %s" % (d, reg, stack, shared, local, ninstr, "  ".join(counts)))
print("\nwhole library: " + "  ".join("%s %d" % kv for kv in tot.items()))
