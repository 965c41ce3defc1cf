"""Turn an `ncu --set full --import-source on` report into the short text summary committed under profiles/:
headline metrics of the captured launch + the SASS lines that collected the most warp-stall samples.
Usage: python profiles/summarize_ncu.py report.ncu-rep > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_active.avg")


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    raw = page(rep, "raw")
    hdr, units, row = raw[0], raw[1], raw[2]
    d = dict(zip(hdr, row))
    u = dict(zip(hdr, units))
    print(f"report: {rep}")
    print(f"kernel: {d.get('Kernel Name')}")
    for k in hdr:
        if any(k.endswith(x) or k == x for x in KEEP) or (k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")):
            try:
                v = float(d[k].replace(",", ""))
            except ValueError:
                continue
            if "stalled" in k and v < 0.3:
                continue
            print(f"  {k} = {d[k]} {u.get(k, '')}")
    src = page(rep, "source")
    h = src[1]
    ia, isrc, iall, iex = h.index("Address"), h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    data = [(int(r[iall] or 0), int(r[iex] or 0), r[ia][-5:], r[isrc].strip()) for r in src[2:] if len(r) > iex]
    tot = sum(x[0] for x in data)
    print(f"warp-stall samples: {tot}; top SASS lines (samples, share, executed, address, instruction):")
    for smp, ex, a, s in sorted(data, key=lambda x: -x[0])[:24]:
        print(f"  {smp:7d} {100 * smp / max(tot, 1):5.1f}%  ex={ex:9d}  {a}  {s[:110]}")


if __name__ == "__main__":
    main(sys.argv[1])
