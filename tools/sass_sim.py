"""Issue-model estimate of a pair loop's FP32-pipe utilisation, from the SASS (no GPU needed).

    python tools/sass_sim.py <file.sass> [kernel filter] [warps per scheduler]

A deliberately small model of one SM sub-partition: `warps` warps run the same innermost pair loop, the
scheduler issues at most one instruction per clock from the first warp (round robin) whose next
instruction has its source registers ready and its pipe free; in order per warp.  Pipes: FP32 (a packed
FFMA2/FMUL2/FADD2 holds it 2 clocks, a scalar FP32 op 1), MUFU (8 clocks per warp instruction: 4 lanes per
clock and sub-partition), ALU (compares, selects, min/max, integer: 2 clocks), LSU.  Latencies: FP32 4,
ALU 4 (+2 for a predicate), MUFU 18, LDS 28.  The number it prints -- FP32 pipe busy clocks / clocks per loop
trip -- is what ncu reports as sm__pipe_fma_cycles_active for a kernel that spends its time in this loop; it
ranks instruction schedules of the same loop (what ptxas made of two versions of the source), which is all it is
used for: the absolute value is within a few points of the measured one for the kernels it was checked on
(profiles/README.md).
"""
import re
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import sass_mix as sm  # noqa: E402

LAT = {"fp32": 4, "mufu": 18, "alu": 4, "lds": 28, "other": 4}


def classify(op):
	if re.match(r"F(FMA|MUL|ADD)2", op):
		return "fp32", 2
	if re.match(r"F(FMA|MUL|ADD)(\.|$)", op):
		return "fp32", 1
	if op.startswith("MUFU"):
		return "mufu", 8
	if op.startswith("LDS"):
		return "lds", 1
	if op.startswith(("BRA", "NOP", "BAR", "WARPSYNC")):
		return "other", 1
	return "alu", 2


def regs_of(op, args):
	"""(dest registers, source registers, dest predicates, source predicates) of one SASS instruction."""
	parts = [a.strip() for a in args.split(",")]
	def expand(tok, width_hint=1):
		m = re.search(r"\bR(\d+)\b", tok)
		if not m:
			return []
		r = int(m.group(1))
		w = 2 if ".F32x2" in tok or ".64" in tok else width_hint
		return list(range(r, r + w))
	def preds(tok):
		return [int(x) for x in re.findall(r"(?<![UR])\bP(\d)\b", tok)]
	dst, src, pd, ps = [], [], [], []
	if not parts:
		return dst, src, pd, ps
	kind, _ = classify(op)
	dwidth = 2 if kind == "fp32" and op[4:5] == "2" or op.endswith("2") and kind == "fp32" else 1
	if op.startswith("LDS"):
		dwidth = {"128": 4, "64": 2}.get(op.split(".")[1] if "." in op else "", 1)
	ndst = 1
	if op.startswith(("FSETP", "ISETP", "UISETP")):
		pd = preds(parts[0]) + preds(parts[1])
		for t in parts[2:]:
			src += expand(t)
			ps += preds(t)
		return dst, src, pd, ps
	if op.startswith(("BRA", "STS", "ST.", "STG", "BAR")):
		ndst = 0
	for i, t in enumerate(parts):
		if i < ndst:
			dst += expand(t, dwidth)
		else:
			src += expand(t, 2 if (kind == "fp32" and dwidth == 2 and ".F32x2" in t) else 1)
			ps += preds(t)
	return dst, src, pd, ps


def simulate(loop, warps=2, trips=6):
	ins = []
	for _, op, args in loop:
		kind, hold = classify(op)
		guard = re.match(r"", "")
		ins.append((kind, hold) + regs_of(op, args))
	n = len(ins)
	ready = [dict() for _ in range(warps)]          # register -> clock its value is available
	pready = [dict() for _ in range(warps)]
	pc = [0] * warps
	done_trips = [0] * warps
	pipe_free = {"fp32": 0, "mufu": 0, "alu": 0, "lds": 0, "other": 0}
	clock, fp_busy, start_clock, start_busy = 0, 0, None, None
	rr = 0
	while min(done_trips) < trips:
		issued = False
		for k in range(warps):
			w = (rr + k) % warps
			if done_trips[w] >= trips:
				continue
			kind, hold, dst, src, pd, ps = ins[pc[w]]
			if pipe_free[kind] > clock:
				continue
			if any(ready[w].get(r, 0) > clock for r in src) or any(pready[w].get(p, 0) > clock for p in ps):
				continue
			# WAW / in-order register reuse is not a hazard in this model
			pipe_free[kind] = clock + hold
			if kind == "fp32":
				fp_busy += hold
			for r in dst:
				ready[w][r] = clock + LAT[kind] + (hold - 1)
			for p in pd:
				pready[w][p] = clock + LAT["alu"] + 2
			pc[w] += 1
			if pc[w] == n:
				pc[w] = 0
				done_trips[w] += 1
				if w == 0 and done_trips[w] == 2:
					start_clock, start_busy = clock, fp_busy
				if w == 0 and done_trips[w] == trips:
					end_clock, end_busy = clock, fp_busy
			rr = (w + 1) % warps
			issued = True
			break
		clock += 1
	return (end_busy - start_busy) / (end_clock - start_clock), (end_clock - start_clock) / (trips - 2)


def main():
	path = Path(sys.argv[1]) if len(sys.argv) > 1 else sm.SASS
	flt = sys.argv[2] if len(sys.argv) > 2 else ""
	warps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
	text = path.read_text()
	import subprocess
	fns = list(sm.functions(text))
	dem = subprocess.run(["c++filt"], input="\n".join(n for n, _ in fns), capture_output=True, text=True).stdout.splitlines()
	for (name, body), d in zip(fns, dem):
		m = re.search(r"m2m_kernel<cvtx::(\w+(?:<\d+>)?), (\d+), (\d+), (\d+)[,>]", d)
		if not m or flt not in d:
			continue
		for lp in sm.inner_loops(body):
			alu = sum(1 for _, op, _ in lp if re.match(r"(FSETP|FSET|FSEL)", op))
			util, clocks = simulate(lp, warps)
			fp = sum(classify(op)[1] for _, op, _ in lp if classify(op)[0] == "fp32")
			print(f"{m.group(1):18s} T={m.group(2)} B={m.group(3)} {'guarded' if alu else 'plain':8s} {len(lp):5d} instr  fp32 clocks {fp:5d}  "
			      f"model: {clocks:8.0f} clocks/trip with {warps} warps -> FP32 pipe {100 * util:5.1f} %")


if __name__ == "__main__":
	main()
