"""Collect measured evidence from gpurun_out/ (scratch) into profiles/ (tracked)."""
import csv
import glob
import json
import os
import re
import shutil
import subprocess

G, P = "gpurun_out", "profiles"
os.makedirs(P, exist_ok=True)
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}


def jl(path):
    try:
        txt = open(path).read().strip().splitlines()
        for line in reversed(txt):
            if line.startswith("{"):
                return json.loads(line)
    except Exception:
        return None
    return None


# ---- headline bench table ------------------------------------------------------------------------------------
rows = []
for f in sorted(glob.glob(f"{G}/ours_n*.json") + glob.glob(f"{G}/ref_n*.json") + glob.glob(f"{G}/bench_*.json")):
    r = jl(f)
    if not r or "value" not in r:
        continue
    rows.append((os.path.basename(f), r))
if rows:
    with open(f"{P}/bench_headline.md", "w") as fh:
        fh.write("# Headline benchmark runs (bench.py, B200, device-timed)\n\n")
        fh.write("| file | impl | model | mode | gpus | global batch | ms/step | samples/s | e2e samples/s | engine | comm | clocks (MHz, reasons) |\n|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for name, r in rows:
            c = r.get("config", {})
            ck = r.get("clocks", {})
            e2e = r.get("e2e", {}).get("value")
            fh.write(f"| {name} | {r.get('impl')} | {c.get('model', '')[:14]} | {c.get('mode')} | {r['n_gpus']} | {c.get('global_batch')} | "
                     f"{r['ms_per_step']:.4f} | {r['value']:.0f} | {e2e and round(e2e)} | {c.get('engine', '')} | {c.get('comm', c.get('backend', ''))} | "
                     f"{ck.get('sm_mhz')} / {ck.get('sm_max_mhz')} {ck.get('reasons')} |\n")
    for name, r in rows:
        json.dump(r, open(f"{P}/{name}", "w"), indent=1)

# ---- launch list -------------------------------------------------------------------------------------------------
for f in sorted(glob.glob(f"{G}/launches*.csv")):
    lines = [ln for ln in open(f) if not ln.startswith("==")]
    recs = list(csv.DictReader(lines))
    if not recs:
        continue
    tail = recs[-40:]
    agg = {}
    for x in tail:
        k = re.sub(r"\(.*", "", x["Kernel Name"])[:60]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(x["Metric Value"])
    with open(f"{P}/{os.path.basename(f).replace('.csv', '.md')}", "w") as fh:
        fh.write("# Kernel launch list of the training step (ncu --metrics gpu__time_duration.sum, cold cache, serialised)\n\n")
        fh.write("Last 40 launches, grouped.  Compare SHARES, not absolutes (B200_PROFILING.md).\n\n| kernel | launches | total ns | mean ns |\n|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write(f"| `{k}` | {n} | {t:.0f} | {t / n:.0f} |\n")

# ---- ncu full report of the GEMM ------------------------------------------------------------------------------------
for rep in sorted(glob.glob(f"{G}/prof_*.ncu-rep")):
    try:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, timeout=120).stdout
    except Exception:
        continue
    rr = list(csv.reader(raw.splitlines()))
    if len(rr) < 3:
        continue
    hdr = rr[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__cycles_active.avg", "gpc__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    with open(f"{P}/{os.path.basename(rep).replace('.ncu-rep', '_ncu.md')}", "w") as fh:
        fh.write(f"# ncu --set full summary ({os.path.basename(rep)})\n\n")
        for row in rr[2:]:
            fh.write("\n".join(f"- {w}: {row[i]} {rr[1][i]}" for w, i in idx) + "\n\n")

# ---- plain copies -----------------------------------------------------------------------------------------------------
for f in ["gemm_bench.json", "gemm_phases.json", "ci_summary.txt", "multi2.log", "multi8.log", "multi4.log"] + \
        [os.path.basename(x) for x in glob.glob(f"{G}/allreduce_sweep_*.json") + glob.glob(f"{G}/kar_profile_*.json")]:
    if os.path.exists(f"{G}/{f}"):
        shutil.copy(f"{G}/{f}", f"{P}/{f}")
if os.path.exists(f"{G}/gemm_bench.json"):
    gb = json.load(open(f"{G}/gemm_bench.json"))
    with open(f"{P}/gemm_vs_cublas.md", "w") as fh:
        fh.write(f"# tcgen05 GEMM vs cuBLAS (bf16, device-timed, inputs rotated through >L2 buffers)\n\nMeasured peak (MEASURED_PEAKS.json): {peaks.get('bf16_tflops')} TFLOP/s burst.\n\n")
        fh.write("| M | N | K | ours BN=128 | ours BN=256 | persistent BN=128 | persistent BN=256 | cuBLAS | best ours / measured peak |\n|---|---|---|---|---|---|---|---|---|\n")
        for r in gb:
            g = lambda k: f"{r[k]:.0f}" if k in r else "-"   # noqa: E731
            fh.write(f"| {r['M']} | {r['N']} | {r['K']} | {g('ours_bn128_tflops')} | {g('ours_bn256_tflops')} | {g('ours_bn128_persistent_tflops')} | "
                     f"{g('ours_bn256_persistent_tflops')} | {r['cublas_tflops']:.0f} | {r['frac_of_measured_peak']:.2f} (of measured) |\n")
print("profiles written:", sorted(os.listdir(P)))
