"""Inner-loop instruction mix of the pair kernels, read from the SASS of the built library (no GPU needed).

    make -C cvortex_b200/csrc sass && python tools/sass_mix.py [filter]

For every `m2m_kernel<Policy, T, B, MINB>` it finds the innermost pair loops (an OPTIMISTIC policy has
two: the plain one and the guarded one that re-evaluates a chain, pair_math.cuh "GUARDS"), counts the
opcodes in each and prints them per (source, target) pair: packed FP32x2 instructions count two
lane-ops, so `lane-ops` is directly comparable with the L column of DESIGN.md section 4, `MUFU` with
S; `ALU` are the compares / selects / min-max (FSETP, FSEL, FMNMX*: 16-lane ALU pipe) and `issue` all
instructions, against the issue bound of one warp instruction per clock and SM sub-partition.
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

SASS = Path(__file__).resolve().parent.parent / "cvortex_b200" / "lib" / "libcvortex.sass"
INS = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);")


def functions(text):
	name, body = None, []
	for line in text.splitlines():
		m = re.search(r"Function : (\S+)", line)
		if m:
			if name:
				yield name, body
			name, body = m.group(1), []
			continue
		m = INS.match(line)
		if m and name:
			body.append((int(m.group(1), 16), m.group(2), m.group(3)))
	if name:
		yield name, body


def inner_loops(body):
	"""Innermost loops (backward branches with no other backward branch inside) that do FP32 work."""
	spans = []
	for addr, op, args in body:
		if not op.startswith("BRA"):
			continue
		m = re.search(r"0x([0-9a-f]+)", args)
		if m and int(m.group(1), 16) < addr:
			spans.append((int(m.group(1), 16), addr))
	out = []
	for lo, hi in spans:
		if any((l2, h2) != (lo, hi) and lo <= l2 and h2 <= hi for l2, h2 in spans):
			continue
		span = [i for i in body if lo <= i[0] <= hi]
		if sum(1 for i in span if re.match(r"F(FMA|MUL|ADD)", i[1])) >= 16 and any(i[1].startswith("LDS") for i in span):
			out.append(span)
	return out


def loop_table(text):
	"""One dict per innermost pair loop of every m2m_kernel instantiation in a cuobjdump -sass dump."""
	fns = list(functions(text))
	dem = subprocess.run(["c++filt"], input="\n".join(n for n, _ in fns), capture_output=True, text=True).stdout.splitlines()
	rows = []
	for (name, body), d in zip(fns, dem):
		m = re.search(r"m2m_kernel<cvtx::(\w+(?:<\d+>)?), (\d+), (\d+), (\d+)[,>]", d)
		if not m:
			continue
		pol, T = m.group(1), int(m.group(2))
		for loop in inner_loops(body):
			c = collections.Counter(op for _, op, _ in loop)
			mufu = sum(v for k, v in c.items() if k.startswith("MUFU"))
			lds = sum(v for k, v in c.items() if k.startswith("LDS"))
			packed = sum(v for k, v in c.items() if re.match(r"F(FMA|MUL|ADD)2", k))
			scalar = sum(v for k, v in c.items() if re.match(r"F(FMA|MUL|ADD)(\.|$)", k))
			per_src = 1 if pol.startswith("P2D") else 2            # LDS.128 per packed source record
			pairs = max(lds // per_src, 1) * T
			alu = sum(v for k, v in c.items() if re.match(r"(FSETP|FSET|FSEL|FMNMX)", k))
			form = "guarded" if alu else "plain"
			if pol.startswith("F3D"):
				# filament tiers (pair_math.cuh FILAMENTS): every fast pair feeds its target's flag through exactly one
				# NaN-keeping minimum -- FMNMX3 in the cancellation-free form, FMNMX in the form that selects per pair
				pairs = max(sum(v for k, v in c.items() if k.startswith("FMNMX") and ".NAN" in k), 1)
				form = "new" if any(k.startswith("FMNMX3") for k in c) else "wide"
			total = len(loop)
			rest = collections.Counter({k: v for k, v in c.items() if not re.match(r"F(FMA|MUL|ADD)", k) and not k.startswith(("MUFU", "LDS"))})
			rows.append({"kernel": d, "policy": pol, "T": T, "form": form, "pairs": pairs,
			             "lane_ops": (packed * 2 + scalar) / pairs, "mufu": mufu / pairs, "alu": alu / pairs, "lds": lds / pairs,
			             "other": (total - packed - scalar - mufu - lds - alu) / pairs, "issue": total / pairs,
			             "tops": ", ".join(f"{k}x{v}" for k, v in rest.most_common(4))})
	return rows


def main():
	flt = sys.argv[1] if len(sys.argv) > 1 else ""
	print(f"{'kernel':26s} {'loop':8s} {'pairs':>5s} {'lane-ops':>8s} {'MUFU':>5s} {'ALU':>5s} {'LDS':>5s} {'other':>6s} {'issue':>6s}   (per pair)   top non-FP32 opcodes")
	for r in loop_table(SASS.read_text()):
		if flt in r["kernel"]:
			print(f"{r['policy'] + ' T=' + str(r['T']):26s} {r['form']:8s} {r['pairs']:5d} {r['lane_ops']:8.2f} {r['mufu']:5.2f} {r['alu']:5.2f} "
			      f"{r['lds']:5.2f} {r['other']:6.2f} {r['issue']:6.2f}   {r['tops']}")


if __name__ == "__main__":
	main()
