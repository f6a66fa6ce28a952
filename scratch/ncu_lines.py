"""Attribute ncu SASS-level samples to CUDA source lines (via nvdisasm -g line markers).
usage: python scratch/ncu_lines.py <rep> <cubin> <kernel-substr> [top]"""
import csv, re, subprocess, sys, collections
rep, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hdr]
ci = {n: h.index(n) for n in ("Source", "# Samples", "Instructions Executed")}
sass = [(r[ci["Source"]].strip(), int(r[ci["# Samples"]] or 0), int(r[ci["Instructions Executed"]] or 0)) for r in rows[hdr + 1:] if len(r) > ci["Instructions Executed"]]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# find function section
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and kname in l]
lines = []
cur = None
infn = False
for l in dis:
    if l.startswith(".text."):
        infn = kname in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
print("sass instrs", len(sass), "disasm instrs", len(lines))
n = min(len(sass), len(lines))
by_line = collections.defaultdict(lambda: [0, 0])
for (ins, smp, ex), ln in zip(sass[:n], lines[:n]):
    by_line[ln][0] += smp
    by_line[ln][1] += ex
tot_s = sum(v[0] for v in by_line.values()); tot_e = sum(v[1] for v in by_line.values())
print("total samples", tot_s, "total warp-instr", tot_e)
src_cache = {}
def src(ln):
    if ln is None: return ""
    f, n = ln
    import glob
    if f not in src_cache:
        c = glob.glob(f"/root/repo/vae_training_b200/csrc/{f}")
        src_cache[f] = open(c[0]).read().splitlines() if c else []
    s = src_cache[f]
    return s[n - 1].strip()[:90] if 0 < n <= len(s) else ""
for ln, (s, e) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{str(ln):28s} samples {100*s/tot_s:5.1f}%  instr {100*e/tot_e:5.1f}%  | {src(ln)}")
# ---- region summary
reg = collections.defaultdict(lambda: [0, 0])
def region(ln):
    import re as _re
    if ln is None: return "none"
    f, n = ln
    def marks(path, pats):
        out = []
        for k, l in enumerate(open(path).read().splitlines(), 1):
            for name, pat in pats:
                if pat in l: out.append((k, name))
        return sorted(out)
    global _mk
    try: _mk
    except NameError:
        _mk = {"common.cuh": marks("/root/repo/vae_training_b200/csrc/common.cuh", [("threefry", "void threefry2x32("), ("threefry", "void threefry2x32_x2("), ("split", "void split2("), ("log1p", "float neg_log1p_neg("), ("erfinv", "float erf_inv_f32("), ("normal", "float normal_from_bits("), ("normal_fast", "float normal_from_bits_fast("), ("det_exp", "float det_expf("), ("warp_sum", "float warp_sum("), ("hosthelpers", "struct AdamHyper")]),
               "fused_small.cu": marks("/root/repo/vae_training_b200/csrc/fused_small.cu", [("init", "---- load state"), ("phaseA", "phase A: random"), ("phaseB", "phase B: one thread"), ("phaseC", "phase C: gradient"), ("store", "---- store state"), ("pre", "namespace vae {")]),
               "fused_mlp.cu": marks("/root/repo/vae_training_b200/csrc/fused_mlp.cu", [("pre", "namespace vae {"), ("barrier", "void cluster_barrier()"), ("slice", "struct Slice"), ("load_rows", "void load_rows("), ("tile_gemm", "void tile_gemm("), ("stage_cols", "void stage_w_cols("), ("stage_rows_t", "void stage_w_rows_t("), ("tile_dw", "void tile_dw("), ("adam", "struct MlpHyper"), ("block_sum", "float block_sum("), ("kernel_init", "fused_mlp_kernel(const"), ("phaseR", "phase R: random"), ("xform", "every CTA: eps rows"), ("enc_fwd", "= encoder forward"), ("reparam", "---- reparameterise"), ("dec_fwd", "= decoder forward"), ("output", "= output, loss"), ("dec_bwd", "= decoder backward"), ("latent", "= latent: d lv"), ("enc_bwd", "= encoder backward"), ("host", "---- host side")])}
    if f in _mk:
        r = "pre"
        for k, name in _mk[f]:
            if n >= k: r = name
        return f.split(".")[0][:6] + ":" + r
    return f
NCTA = int(__import__("os").environ.get("NCTA", "888")); NSTEP = int(__import__("os").environ.get("NSTEP", "2000"))
for ln, (s, e) in by_line.items():
    r = region(ln); reg[r][0] += s; reg[r][1] += e
for r, (s, e) in sorted(reg.items(), key=lambda kv: -kv[1][1]):
    print(f"{r:22s} samples {100*s/tot_s:5.1f}%  instr {100*e/tot_e:5.1f}%  warp-instr/CTA-step {e/(NCTA*NSTEP):8.1f}")
