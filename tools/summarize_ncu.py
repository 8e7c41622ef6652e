"""Turn ncu outputs from gpurun_out/ into the tracked summaries under profiles/.
usage: python tools/summarize_ncu.py <tag> <launch_list.csv> <full_report.ncu-rep> [<full_report2.ncu-rep> ...]"""
import collections, csv, json, re, subprocess, sys, os


def short(name):
    """'void adb::fast_cells_kernel<64>(...)' -> 'fast_cells_kernel'"""
    return re.sub(r'<.*', '', name.split('(')[0].replace('void ', '').replace('adb::', '')).strip()


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem']

def launch_shares(path, out, note):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]; ki, vi = h.index('Kernel Name'), h.index('Metric Value')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        try: v = float(r[vi].replace(',', ''))
        except ValueError: continue
        n = short(r[ki])[:90]; agg[n][0] += 1; agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    with open(out, 'w') as f:
        f.write(f"# {note}\n# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\nkernel,launches,total_us,share\n")
        for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"\"{n}\",{c},{t / 1e3:.1f},{t / tot:.4f}\n")

def full_summary(reps, out_csv, out_json, note):
    lines = [f"# {note}", "# ncu --set full --clock-control none --import-source on; one row per profiled launch", "kernel," + ",".join(WANT)]
    traffic = {}
    for rep in reps:
        raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        h = rows[0]; ki = h.index('Kernel Name')
        idx = [h.index(w) if w in h else None for w in WANT]
        for r in rows[2:]:
            name = short(r[ki])
            vals = [r[i] if i is not None else "" for i in idx]
            lines.append('"%s",' % name + ",".join(v.replace(',', '') for v in vals))
            try:
                units = [rows[1][i] for i in idx[1:3]]
                conv = lambda v, u: float(v.replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
                t = conv(vals[1], units[0]) + conv(vals[2], units[1])
                traffic.setdefault(name, []).append({"dram_bytes": t, "duration_us": float(vals[0].replace(',', '')), "grid": vals[WANT.index('launch__grid_size')]})
            except Exception:
                pass
    open(out_csv, 'w').write("\n".join(lines) + "\n")
    json.dump(traffic, open(out_json, 'w'), indent=1)

if __name__ == "__main__":
    tag, ll, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    os.makedirs("profiles", exist_ok=True)
    launch_shares(ll, f"profiles/{tag}_launch_shares.csv", f"{tag}: launch list of `python bench.py --steps 2 --warmup 3 --pairs 128 --no-cpu-baseline` (ORB step + BA + search sections)")
    full_summary(reps, f"profiles/{tag}_ncu_full_summary.csv", f"profiles/{tag}_dram_traffic.json", f"{tag}: bench.py --pairs 128 (128 frames per extractor launch), bench_ba config 4, bench_search (128 frames)")
