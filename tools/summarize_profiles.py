"""Turns ncu outputs in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py <round tag> [launches.csv] [prof.ncu-rep]

Writes profiles/<tag>_launches.md (per-kernel share of a bench step, from the gpu__time_duration pass),
profiles/<tag>_ncu_full.md (key metrics of the --set full capture) and refreshes profiles/ncu_summary.json
(dram bytes per launch of the dominant kernels, read by bench.py for roofline.traffic)."""
import csv, io, json, os, subprocess, sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
launches = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "launches.csv")
rep = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "gpurun_out", "prof_step.ncu-rep")
RAYS = int(os.environ.get("PROFILE_RAYS", 5529600))  # rays per launch of the --set full capture (8 views 960x720)
sys.path.insert(0, ROOT)
from drt_b200 import build as _build  # noqa: E402
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)


def short(name):
    n = name.split("(")[0]
    for p in ("void ", "drt::", "at::native::", "cub::CUB_300001_SM_1000::", "cub::"):
        n = n.replace(p, "")
    return n[:70]


if os.path.exists(launches):
    lines = [l for l in open(launches) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    per = defaultdict(lambda: [0, 0.0])
    order = []
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", ""))
        u = r.get("Metric Unit", "ns")
        v_us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        if k not in per:
            order.append(k)
        per[k][0] += 1
        per[k][1] += v_us
    tot = sum(v[1] for v in per.values())
    with open(os.path.join(out_dir, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: per-kernel device time, `ncu --metrics gpu__time_duration.sum --clock-control none` over\n"
                "`python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline` with `--profile-from-start off` (bench.py brackets\n"
                "exactly the K timed steps with cudaProfilerStart/Stop): C4, 72 views, 2 steps.\n"
                "Times are cold-cache and serialised: compare SHARES, not absolutes.\n\n"
                "| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k in sorted(per, key=lambda k: -per[k][1]):
            f.write(f"| `{k}` | {per[k][0]} | {per[k][1]:.1f} | {100 * per[k][1] / tot:.1f} % |\n")
    print("wrote launches summary:", len(per), "kernels, total", round(tot / 1e3, 2), "ms")

if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    keys = OrderedDict([
        ("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("launch__registers_per_thread", "registers"), ("smsp__inst_executed.sum", "warp instructions"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__thread_inst_executed.sum", "lane instructions"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe % of peak"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe % of peak"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe % of peak"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data-pipe wavefronts % of peak"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard / issue"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: fixed-latency wait / issue"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle / issue"),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected / issue"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving / issue"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier / issue"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: LG throttle / issue"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction / issue")])
    summ = {}
    with open(os.path.join(out_dir, f"{tag}_ncu_full.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` of one bench step\n"
                f"(`python bench.py --steps 1 --warmup 3 [--views V] --graph off` under `--profile-from-start off`; C4 mesh, {RAYS:,} rays per launch;\n"
                f"kernel sources hash {_build.source_hash()}).\n\n")
        for r in rows[2:]:
            name = short(r[hdr.index("Kernel Name")])
            f.write(f"## `{name}`\n\n| metric | value |\n|---|---:|\n")
            d = {}
            for k, lab in keys.items():
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {lab} (`{k}`) | {r[i]} {units[i]} |\n")
                    d[k] = (r[i], units[i])
            f.write("\n")

            def to_bytes(v, u):
                v = float(v.replace(",", ""))
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            def num(k):
                try:
                    return float(d[k][0].replace(",", ""))
                except Exception:
                    return None
            if "dram__bytes_read.sum" in d:
                e = summ.setdefault(name, {"launches": 0, "dram_bytes_per_launch": 0.0, "warp_inst": 0.0, "lane_inst": 0.0, "duration_us": 0.0,
                                           "rays_per_launch": RAYS})
                e["launches"] += 1
                e["dram_bytes_per_launch"] += to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"])
                e["warp_inst"] += num("smsp__inst_executed.sum") or 0.0
                e["lane_inst"] += (num("smsp__thread_inst_executed.sum") or
                                   (num("smsp__inst_executed.sum") or 0.0) * (num("smsp__thread_inst_executed_per_inst_executed.ratio") or 0.0))
                dur, du = d["gpu__time_duration.sum"]
                e["duration_us"] += float(dur.replace(",", "")) * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(du, 1.0)
                e["issue_active_pct"] = num("smsp__issue_active.avg.pct_of_peak_sustained_active")
    js = os.path.join(out_dir, "ncu_summary.json")
    allj = json.load(open(js)) if os.path.exists(js) else {}
    allj[tag] = summ
    # bench.py reads the forward group: sum of the five wavefront launches, scaled per ray
    fwd = [v for k, v in summ.items() if k.startswith("wf_")]
    if fwd:
        allj["trace_fwd"] = {"dram_bytes_per_ray": sum(v["dram_bytes_per_launch"] for v in fwd) / RAYS, "from": tag, "source_hash": _build.source_hash()}
    # the forward stages inside drt_ray_loss_step (ls_beam/q1_tiles/r1/q2/r2/q3; ls_loss_bwd is the backward)
    ls = [v for k, v in summ.items() if k.startswith("ls_") and not k.startswith("ls_loss_bwd")]
    if ls:
        wi, li = sum(v["warp_inst"] for v in ls), sum(v["lane_inst"] for v in ls)
        allj["loss_step_fwd"] = {"dram_bytes_per_ray": sum(v["dram_bytes_per_launch"] for v in ls) / RAYS, "from": tag,
                                 "source_hash": _build.source_hash(), "rays_per_launch": RAYS,
                                 "warp_inst_per_ray": wi / RAYS, "lane_inst_per_ray": li / RAYS, "lanes_per_inst": li / wi if wi else None,
                                 "kernel_us_under_ncu": sum(v["duration_us"] for v in ls)}
    json.dump(allj, open(js, "w"), indent=1)
    print("wrote ncu full summary:", list(summ))
