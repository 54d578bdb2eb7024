#!/usr/bin/env python3
"""What the hot kernels compile to (run here after a build; writes profiles/r03_sass_*.txt):
   registers / stack / spills from the ptxas log, and the SASS instruction mix of the kernel bodies - 128-bit global loads of the
   node / leaf records, shared-memory traffic of the pools, no local loads/stores outside the cold walk_subtree function, the
   TMA instructions of the MLAA strip kernel."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = os.path.join(ROOT, "renderer_b200", "build")


def functions(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
    cur, d = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1); d[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4,5}\*/", line):
            d[cur].append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).rstrip())
    return d


def ptxas(log, needle):
    lines = open(log).read().splitlines()
    res = []
    for i, l in enumerate(lines):
        if "Function properties for" in l and needle in l:
            res.append(lines[i + 1].strip() + " | " + lines[i + 2].strip().replace("ptxas info    : ", ""))
    return res


def mix(body):
    c = collections.Counter()
    for l in body:
        t = l.split("*/", 1)[1].split()
        op = t[1] if t[0].startswith("@") else t[0]
        c[op.rstrip(";")] += 1
    return c


def report(path, title, obj, log, needle, want, excerpt=None):
    f = functions(obj)
    with open(path, "w") as o:
        o.write(f"# {title}\n# cuobjdump -sass {os.path.relpath(obj, ROOT)} (sm_100a), nvcc 12.9; produced by tools/sass_evidence.py\n\n")
        for name, body in f.items():
            if needle not in name:
                continue
            c = mix(body)
            short = subprocess.run(["c++filt", name], stdout=subprocess.PIPE, text=True).stdout.strip()[:200]
            o.write(f"## {short}\n")
            for p in ptxas(log, name):
                o.write(f"ptxas: {p}\n")
            o.write(f"SASS instructions: {len(body)}\n")
            o.write("instruction mix (static): " + ", ".join(f"{k} {v}" for k, v in c.most_common() if any(k.startswith(w) for w in want)) + "\n")
            loc = [l for l in body if re.search(r"\b(LDL|STL)\b", l)]
            hot = [l for l in body if re.search(r"REDUX|VOTE", l)]
            last_hot = hot[-1].split("*/")[0].split("/*")[1] if hot else "-"
            o.write(f"local-memory instructions (LDL/STL): {len(loc)}" + (f"; the pass loop (its last REDUX/VOTE) ends at 0x{last_hot}; LDL/STL at: " +
                    " ".join("0x" + l.split("*/")[0].split("/*")[1] for l in loc) if loc else "") + "\n")
            if excerpt:
                for l in body:
                    if re.search(excerpt, l):
                        o.write("    " + l.strip() + "\n")
            o.write("\n")


os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
report(os.path.join(ROOT, "profiles", "r03_sass_rt_pool_kernel.txt"),
       "rt_pool_kernel (all instantiations). walk_subtree (__noinline__, the cold overflow path) is emitted behind the kernel body: its private\n"
       "# DFS stack is the local memory at the high addresses; the 64-register instantiation (4 CTAs per SM) additionally spills 24 bytes;\n"
       "# node records (4 x LDG.E.128) and leaf records (5 x LDG.E.128) are fetched with 128-bit loads, slot state with LDS.128, pool entries with LDS.64/STS.64",
       os.path.join(B, "rt_pool.cu.o"), os.path.join(B, "rt_pool.cu.o.log"), "rt_pool",
       ("LDG", "LDS", "STS", "ATOMS", "VOTE", "REDUX", "LDL", "STL", "FFMA", "FMNMX", "CALL", "WARPSYNC", "STG"), r"LDG\.E\.128|ATOMS|REDUX")
report(os.path.join(ROOT, "profiles", "r03_sass_mlaa_tma.txt"),
       "mlaa_blend_vstrip_tma_kernel: the strip is loaded with UTMALDG.2D (cp.async.bulk.tensor.2d, completion on an mbarrier: SYNCS.*) and\n"
       "# written back with UTMASTG.2D + UTMACMDFLUSH (bulk_group commit / wait)",
       os.path.join(B, "mlaa_kernels.cu.o"), os.path.join(B, "mlaa_kernels.cu.o.log"), "vstrip_tma",
       ("UTMA", "SYNCS", "LDS", "STS", "FENCE", "BAR"), r"UTMA|SYNCS|FENCE")
print("written: profiles/r03_sass_rt_pool_kernel.txt, profiles/r03_sass_mlaa_tma.txt")
