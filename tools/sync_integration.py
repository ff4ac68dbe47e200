#!/usr/bin/env python
"""Rewrite the block of INTEGRATION.md between the ref_binding markers with examples/ref_binding.py (verbatim)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
doc_path = os.path.join(ROOT, "INTEGRATION.md")
doc = open(doc_path).read()
src = open(os.path.join(ROOT, "examples", "ref_binding.py")).read().rstrip("\n")
block = "<!-- ref_binding:begin -->\n```python\n" + src + "\n```\n<!-- ref_binding:end -->"
new, n = re.subn(r"<!-- ref_binding:begin -->.*?<!-- ref_binding:end -->", lambda m: block, doc, flags=re.S)
assert n == 1, "markers not found in INTEGRATION.md"
open(doc_path, "w").write(new)
print("INTEGRATION.md updated")
