"""Extract the literal node examples of the reference's docs/quickstart.rst (the `::` blocks that hold a `Node { ... }`) into
tests/golden/quickstart_nodes.json — reference-AUTHORED .vnf text for the parser tests (the reference ships no scene files).

    python tools/extract_quickstart.py [/root/reference/docs/quickstart.rst]

Each entry: {"line": first line of the block in the .rst, "text": the block with its common indentation removed}."""
import json, os, re, sys

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/docs/quickstart.rst"
lines = open(src).read().split("\n")
out = []
i = 0
while i < len(lines):
    if lines[i].rstrip().endswith("::"):
        j = i + 1
        while j < len(lines) and lines[j].strip() == "":
            j += 1
        blk = []
        first = j
        while j < len(lines) and (lines[j].strip() == "" or lines[j][:1] in (" ", "\t")):
            blk.append(lines[j])
            j += 1
        while blk and blk[-1].strip() == "":
            blk.pop()
        text = "\n".join(blk)
        if re.search(r"^\s*\w+\s*\{", text, re.M):
            out.append({"line": first + 1, "text": text})
        i = j
    else:
        i += 1
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "quickstart_nodes.json")
json.dump({"source": "docs/quickstart.rst of jamiec7919/vermeer (reference snapshot)", "blocks": out}, open(dst, "w"), indent=1)
print(len(out), "blocks ->", dst)
