"""Share of each kernel inside ONE evaluator step (`value` arm of bench.py) from an ncu launch list
(`--metrics gpu__time_duration.sum`): a step = k_prep_rows* .. k_ap_finalize around a full-size k_dist_tc launch.
The launch list of a whole bench run also holds the chunked e2e launches and the re-ranking kernels, which is why the
per-kernel averages of the plain summary are not per-step figures.

    python scripts/step_share.py profiles/r02/launches_bench.csv [min_gemm_us]
"""
import csv, re, sys


def main(path, min_gemm_us=4000.0):
    with open(path) as f:
        rows = list(csv.reader(l for l in f if l.startswith('"')))[1:]
    L = [(re.sub(r"^void ", "", r[4]).split("(")[0], float(r[-1]) / 1e3) for r in rows]
    steps = []
    for i, (k, us) in enumerate(L):
        if "k_dist_tc" in k and us >= min_gemm_us and i > 0 and "k_prep_rows" in L[i - 1][0]:
            j = i + 1
            while j < len(L) and "k_ap_finalize" not in L[j][0] and j - i < 12:
                j += 1
            if j < len(L) and "k_ap_finalize" in L[j][0]:
                steps.append(L[i - 1: j + 1])
    if not steps:
        print("no evaluator step found")
        return
    print(f"# {len(steps)} evaluator steps (prep -> full-size distance GEMM -> label kernels -> k_rank_count -> k_ap_finalize) in {path}")
    names = [k for k, _ in steps[0]]
    tot = sum(sum(us for _, us in s) for s in steps) / len(steps)
    print(f"{'kernel':60s} {'avg us':>9s} {'share':>7s}")
    for idx, name in enumerate(names):
        avg = sum(s[idx][1] for s in steps if len(s) == len(names)) / sum(1 for s in steps if len(s) == len(names))
        print(f"{name[:60]:60s} {avg:9.1f} {100 * avg / tot:6.1f}%")
    print(f"{'step total (serialised, cold cache)':60s} {tot:9.1f}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 4000.0)
