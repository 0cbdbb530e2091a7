# SASS / ptxas evidence for the hot kernels (runs on the build box, no GPU): instruction census per kernel, resource usage, and the
# tcgen05 issue section of the skipping product kernel.  usage: bash tools/sass_evidence.sh
set -e
LIB=gaussian_process_transportation_b200/lib/libgptb200.so
OUT=profiles
cuobjdump -sass $LIB > /tmp/all.sass
python - <<'PY'
import re, collections
txt = open('/tmp/all.sass').read()
funcs = re.split(r"\n\s*Function : ", txt)[1:]
want = ["UTCIMMA", "LDTM", "UTMALDG", "UTCBAR", "DMMA", "DFMA", "SYNCS", "ELECT", "LDGSTS", "HMMA"]
rows = []
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    c = collections.Counter(m for m in re.findall(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, flags=re.M))
    agg = {w: sum(v for k, v in c.items() if k.split(".")[0] == w) for w in want}
    rows.append((name, sum(c.values()), agg))
hot = ["ozaki_trmm_kernel", "trmm_sumsq_kernel", "kstar_kernel", "potrf_trailing", "potrf_panel", "potrf_diag", "potrf_spine", "trsv_back_chain", "trtri_level", "kinv_kernel", "gram_lower", "lml_grad_kernel", "slice_rows", "finalize_kernel"]
with open('profiles/r02_sass_census.txt', 'w') as o:
    o.write("# cuobjdump -sass gaussian_process_transportation_b200/lib/libgptb200.so: instruction census per kernel (tools/sass_evidence.sh)\n")
    o.write("# UTCIMMA = tcgen05.mma kind::i8, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit, DMMA = FP64 mma.sync, SYNCS = mbarrier ops\n")
    tot = collections.Counter()
    for name, n, agg in rows:
        for k, v in agg.items(): tot[k] += v
        if any(h in name for h in hot):
            o.write(f"{name}\n    instructions {n}  " + "  ".join(f"{k} {v}" for k, v in agg.items() if v) + "\n")
    o.write("TOTAL over all kernels: " + "  ".join(f"{k} {v}" for k, v in tot.items() if v) + "\n")
print(open('profiles/r02_sass_census.txt').read()[-400:])
PY
# the tcgen05 issue section of ozaki_trmm_kernel<5, true> (one case of the (za, zb) dispatch) and the dense DMMA inner loop
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC --split-compile 0 -Xptxas -v -o /tmp/lib_v.so gaussian_process_transportation_b200/csrc/gptb200.cu 2> /tmp/ptxas_v.txt || true
python - <<'PY'
import re
t = open('/tmp/ptxas_v.txt').read()
blocks = re.findall(r"Compiling entry function '([^']+)' for 'sm_100a'\n(?:ptxas info\s+: Function properties for [^\n]+\n\s+([^\n]+)\n)?ptxas info\s+: Used ([^\n]+)", t)
hot = ["ozaki_trmm_kernelILi5", "ozaki_trmm_kernelILi6", "trmm_sumsq", "kstar_kernelILi3ELi3ELi2ELi5ELi8", "kstar_kernelILi3ELi3ELi0", "potrf_trailing64", "potrf_diag", "potrf_spine", "trsv_back_chain", "potrf_panel", "lml_grad_kernelILi3", "gram_lower_kernelILi3", "finalize_kernelILi3ELi3"]
with open('profiles/r02_ptxas_v.txt', 'w') as o:
    o.write("# nvcc -Xptxas -v (sm_100a) resource usage of the hot kernels (tools/sass_evidence.sh); stack / spill line shown where non-zero\n")
    spills = re.findall(r"(\d+) bytes spill stores", t)
    o.write(f"# spill stores over the whole library: max {max(map(int, spills)) if spills else 0} bytes\n")
    for name, props, used in blocks:
        if any(h in name for h in hot):
            o.write(f"{name}\n    {used}\n")
            if props and not props.strip().startswith("0 bytes stack frame, 0 bytes spill stores"):
                o.write(f"    {props.strip()}\n")
print(open('profiles/r02_ptxas_v.txt').read()[:1500])
PY
python - <<'PY'
import re
txt = open('/tmp/all.sass').read()
funcs = re.split(r"\n\s*Function : ", txt)[1:]
f = [x for x in funcs if x.startswith("_ZN4gptb2oz17ozaki_trmm_kernelILi5ELb1")][0]
lines = [l.strip()[:120] for l in f.split("\n") if not re.match(r"^\s*/\* 0x", l)]
i0 = [i for i, l in enumerate(lines) if "UTCIMMA" in l][0]
open('profiles/r02_sass_ozaki_issue_excerpt.txt', 'w').write(
    "# ozaki_trmm_kernel<5, true> (cuobjdump -sass): first specialised issue block -- wide UTCIMMA (tcgen05.mma kind::i8) products of one\n"
    "# 64-byte chunk, followed by its UTCBAR (tcgen05.commit); UTMALDG = the producer's TMA loads\n" + "\n".join(lines[i0 - 45:i0 + 45]) + "\n")
PY
