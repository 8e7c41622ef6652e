"""Turn ncu outputs from gpurun_out/ into the tracked summaries under profiles/.
usage: python tools/summarize_ncu.py <tag> <launch_list.csv> <raw1.csv|.ncu-rep> [...]
Every value is converted to a fixed unit that is part of the column name (bytes, us, %, count), whatever unit ncu chose per row."""
import collections, csv, json, re, subprocess, sys, os


def short(name):
    """'void adb::fast_cells_kernel<64>(...)' -> 'fast_cells_kernel'"""
    return re.sub(r'<.*', '', name.split('(')[0].replace('void ', '').replace('adb::', '')).strip()


# metric -> (column name, unit family)
WANT = [('gpu__time_duration.sum', 'duration_us', 'time'), ('dram__bytes_read.sum', 'dram_read_bytes', 'bytes'), ('dram__bytes_write.sum', 'dram_write_bytes', 'bytes'),
        ('lts__t_bytes.sum', 'l2_bytes', 'bytes'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm_throughput_pct', 'pct'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram_throughput_pct', 'pct'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps_active_pct', 'pct'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue_active_pct', 'pct'),
        ('smsp__inst_executed.sum', 'warp_instructions', 'count'),
        ('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'pipe_alu_pct', 'pct'),
        ('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'pipe_fma_pct', 'pct'),
        ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'pipe_fp64_pct', 'pct'),
        ('sm__inst_executed_pipe_fp64.sum', 'inst_pipe_fp64', 'count'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'pipe_tensor_pct_of_active', 'pct'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'pipe_tensor_pct_of_elapsed_all_sms', 'pct'),
        ('sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active', 'pipe_tensor_dmma_inst_pct_of_active', 'pct'),
        ('sm__inst_executed_pipe_tensor.sum', 'inst_pipe_tensor', 'count'),
        ('sm__cycles_active.avg', 'sm_cycles_active_avg', 'count'), ('sm__cycles_elapsed.avg', 'sm_cycles_elapsed_avg', 'count'),
        ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'pipe_lsu_pct', 'pct'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem_bank_conflicts', 'count'),
        ('launch__registers_per_thread', 'registers', 'count'), ('launch__grid_size', 'grid', 'count'), ('launch__block_size', 'block', 'count'),
        ('launch__cluster_size', 'cluster', 'count'), ('launch__waves_per_multiprocessor', 'waves_per_sm', 'count'),
        ('launch__occupancy_limit_registers', 'occ_limit_regs', 'count'), ('launch__occupancy_limit_shared_mem', 'occ_limit_smem', 'count')]
SCALE = {'bytes': {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12},
         'time': {'ns': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3, 'second': 1e6, 's': 1e6}}


def norm(v, unit, fam):
    try:
        x = float(v.replace(',', ''))
    except ValueError:
        return ""
    if fam in SCALE:
        x *= SCALE[fam].get(unit, 1)
    return ("%.6g" % x)


def launch_shares(path, out, note):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]; ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = norm(r[vi], r[ui], 'time')
        if v == "":
            continue
        n = short(r[ki])[:90]; agg[n][0] += 1; agg[n][1] += float(v)
    tot = sum(v[1] for v in agg.values())
    with open(out, 'w') as f:
        f.write(f"# {note}\n# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\nkernel,launches,total_us,share\n")
        for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"\"{n}\",{c},{t:.1f},{t / tot:.4f}\n")


def full_summary(reps, out_csv, out_json, note):
    lines = [f"# {note}", "# ncu --set full --clock-control none (+ tensor / fp64 pipe counters); one row per profiled launch; units are in the column names",
             "kernel," + ",".join(c for _, c, _ in WANT)]
    traffic = {}
    for rep in reps:
        raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = [r for r in csv.reader(raw.splitlines()) if len(r) > 5]
        if len(rows) < 3:
            continue
        h = rows[0]; ki = h.index('Kernel Name'); units = rows[1]
        idx = [h.index(m) if m in h else None for m, _, _ in WANT]
        for r in rows[2:]:
            name = short(r[ki])
            vals = [norm(r[i], units[i], fam) if i is not None else "" for i, (_, _, fam) in zip(idx, WANT)]
            lines.append('"%s",' % name + ",".join(vals))
            try:
                traffic.setdefault(name, []).append({"dram_bytes": float(vals[1]) + float(vals[2]), "duration_us": float(vals[0]), "grid": vals[[c for _, c, _ in WANT].index('grid')]})
            except Exception:
                pass
    open(out_csv, 'w').write("\n".join(lines) + "\n")
    json.dump(traffic, open(out_json, 'w'), indent=1)


if __name__ == "__main__":
    tag, ll, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    os.makedirs("profiles", exist_ok=True)
    note = os.environ.get("NCU_NOTE", f"{tag}")
    if os.path.exists(ll):
        launch_shares(ll, f"profiles/{tag}_launch_shares.csv", f"{tag}: launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (default 3072 pairs per step; ORB step + masked + BA + search sections)")
    full_summary(reps, f"profiles/{tag}_ncu_full_summary.csv", f"profiles/{tag}_dram_traffic.json", note)
