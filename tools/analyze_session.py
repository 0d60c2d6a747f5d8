#!/usr/bin/env python3
"""Turns the outputs of tools/gpu_session.sh (gpurun_out/) into one markdown summary for profiles/:

   python tools/analyze_session.py [gpurun_out] > profiles/rNN_session_summary.md

Reads whatever is there: bench_n*.json, configs.jsonl, sweep_*.jsonl, sweep_ntt.jsonl, sharded_ntt_n*.json,
plonk_n*.json, session.log (test tallies).  Purely a formatter: no numbers are computed beyond ratios to the
default configuration of the same sweep."""
import glob
import json
import os
import sys

D = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"


def jl(path):
    out = []
    try:
        for ln in open(path):
            ln = ln.strip()
            if ln.startswith("{"):
                try:
                    out.append(json.loads(ln))
                except Exception:
                    pass
    except FileNotFoundError:
        pass
    return out


def main():
    print(f"# GPU session summary ({D})\n")
    log = os.path.join(D, "session.log")
    if os.path.exists(log):
        print("## Parity suites\n")
        for ln in open(log):
            if " passed" in ln or " failed" in ln or "xfailed" in ln or "xpassed" in ln or ln.startswith(("XPASS", "XFAIL", "FAILED")):
                print("    " + ln.rstrip())
        print()
    for path in sorted(glob.glob(os.path.join(D, "bench_n*.json"))):
        for r in jl(path):
            print(f"## {os.path.basename(path)}\n")
            print(f"* value {r.get('value'):.4g} {r.get('unit')} at {r.get('n_gpus')} GPU(s), {r.get('ms_per_step'):.3f} ms/step; "
                  f"e2e {r.get('e2e', {}).get('value', float('nan')):.4g}; e2e_submit {(r.get('e2e_submit') or {}).get('value')}")
            print(f"* stages (ms): {r.get('stage_ms')}")
            rf = r.get("roofline", {})
            print(f"* roofline: {rf.get('achieved', float('nan')):.1f} / {rf.get('peak')} GB/s = {rf.get('frac', float('nan')):.3f}")
            print(f"* clocks: {r.get('clocks')}")
            print(f"* cpu_baseline: {r.get('cpu_baseline')}")
            print(f"* groth16: {r.get('groth16')}\n")
    sweeps = sorted(glob.glob(os.path.join(D, "sweep_*.jsonl")))
    if sweeps:
        print("## MSM / NTT sweeps\n")
    for path in sweeps:
        rows = jl(path)
        if not rows:
            continue
        print(f"### {os.path.basename(path)}\n")
        if "tile_log" in rows[0]:
            print("| curve | log2n | radix8 | tile_log | ms | matches default |\n|---|---|---|---|---|---|")
            for r in rows:
                print(f"| {r['curve']} | {r['log2n']} | {r.get('radix8', 0)} | {r['tile_log']} | {r['ms']:.4f} | {r['matches_default']} |")
            print()
            continue
        knobs = [k for k in rows[0] if k.startswith("GB200_") or k == "lib"]
        print("| curve | G | log2n | " + " | ".join(knobs) + " | correct | standalone ms | pipelined ms | vs first | accumulate ms | offsets+levels ms |")
        print("|---|---|---|" + "---|" * len(knobs) + "---|---|---|---|---|---|")
        base = {}
        for r in rows:
            key = (r.get("curve"), r.get("group"), r.get("log2n"), r.get("lib"))
            if "error" in r:
                print(f"| {r.get('curve')} | {r.get('group')} | {r.get('log2n')} | " + " | ".join(str(r.get(k)) for k in knobs) + f" | ERROR {r['error'][:60]} |")
                continue
            base.setdefault(key, r["ms_pipelined"])
            st = r.get("stage_ms", {})
            print(f"| {r['curve']} | {r['group']} | {r['log2n']} | " + " | ".join(str(r.get(k)) for k in knobs)
                  + f" | {r['correct']} | {r['ms_standalone']:.3f} | {r['ms_pipelined']:.3f} | {r['ms_pipelined'] / base[key]:.3f} | "
                  f"{st.get('accumulate')} | {st.get('offsets_scan')} |")
        print()
    # compile-time variants against the default library on the same configuration, best first
    rows = jl(os.path.join(D, "sweep_optlib.jsonl"))
    if rows:
        print("## Compile-time variants vs the default library (sweep_optlib.jsonl), best first per configuration\n")
        print("| curve | G | log2n | lib | correct | pipelined ms | vs default | accumulate ms | vs default |\n|---|---|---|---|---|---|---|---|---|")
        by = {}
        for r in rows:
            if "error" not in r:
                by.setdefault((r["curve"], r["group"], r["log2n"]), []).append(r)
        for key, rs in sorted(by.items()):
            d = next((r for r in rs if r.get("lib") == "default"), None)
            for r in sorted(rs, key=lambda r: r["ms_pipelined"]):
                acc = r.get("stage_ms", {}).get("accumulate")
                rel = f"{r['ms_pipelined'] / d['ms_pipelined']:.3f}" if d else "n/a"
                dacc = d.get("stage_ms", {}).get("accumulate") if d else None
                rel_acc = f"{acc / dacc:.3f}" if (acc and dacc) else "n/a"
                print(f"| {key[0]} | {key[1]} | {key[2]} | {r.get('lib')} | {r['correct']} | {r['ms_pipelined']:.3f} | {rel} | {acc} | {rel_acc} |")
        print()
    for pat, title in (("configs.jsonl", "Other configurations"), ("sharded_ntt_n*.json", "Sharded NTT"), ("plonk_n*.json", "Sharded PLONK")):
        files = sorted(glob.glob(os.path.join(D, pat)))
        if not files:
            continue
        print(f"## {title}\n")
        for path in files:
            for r in jl(path):
                short = {k: v for k, v in r.items() if k not in ("stage_ms_rank0", "note", "includes", "excludes", "data")}
                print(f"* `{os.path.basename(path)}`: {json.dumps(short)[:900]}")
        print()


if __name__ == "__main__":
    main()
