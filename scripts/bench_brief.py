"""Print the key numbers of a bench.py JSON line (stdin or file argument)."""
import json, sys
txt = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
line = [l for l in txt.splitlines() if l.startswith("{")][-1]
d = json.loads(line)
tag = sys.argv[2] if len(sys.argv) > 2 else ""
print(tag, "headline %.1f Mrays/s (%.3f ms/step), e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]),
      "| cpu", None if not d.get("cpu_baseline") else round(d["cpu_baseline"]["value"], 2))
for k, c in (d.get("configs") or {}).items():
    if not c:
        continue
    s = c["stage_ms_per_step"]
    print("  %s: %.1f Mrays/s %.2f ms/step e2e %.1f (%.2f ms) | closest %.2f shadow %.2f shade %.2f other %.2f | cpu %s" % (
        k, c["value"], c["ms_per_step"], c["e2e"]["value"], c["e2e"]["ms_per_step"], s["closest_traversal"], s["shadow_traversal"], s["shading"],
        s["raygen_resolve_accumulate"], None if not c.get("cpu_baseline") else round(c["cpu_baseline"]["value"], 2)))
w = d.get("incoherent_wavefront")
if w:
    print("  wavefront: %.1f Mrays/s, hit %.2f, nodesT %.1f trisT %.1f, e2e %s, cpu %s" % (w["value"], w["hit_fraction"], w["nodesT_per_ray"], w["trisT_per_ray"], w["e2e"], (w.get("cpu") or {}).get("value")))
w = d.get("incoherent_diffuse")
if w:
    print("  diffuse 4-bounce (C3): %.1f Mrays/s, hit %.2f, nodesT %.1f trisT %.1f, e2e %s, cpu %s | per bounce %s" % (w["value"], w["hit_fraction"], w["nodesT_per_ray"], w["trisT_per_ray"], w["e2e"], (w.get("cpu") or {}).get("value"),
          [(b["bounce"], b["rays"], round(b["value"]), round(b["hit_fraction"], 2)) for b in w["per_bounce"]]))
