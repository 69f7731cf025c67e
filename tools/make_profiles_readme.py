#!/usr/bin/env python3
"""Regenerates the round-2 tables of profiles/README.md from the bench JSON lines committed under profiles/
(tools/gpu/profiles_table.py makes the rows); the prose between the tables and the round-1 section are kept."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(ROOT, "profiles", "README.md")
text = open(path).read()


def table(prefix):
    return subprocess.run(["python", os.path.join(ROOT, "tools", "gpu", "profiles_table.py"), prefix], capture_output=True,
                          text=True).stdout.strip()


def replace_table_after(text, heading, new_table):
    i = text.index(heading)
    j = text.index("| file |", i)
    k = j
    while k < len(text) and text[k:k + 1] == "|":
        k = text.index("\n", k) + 1
    return text[:j] + new_table + "\n" + text[k:]


text = replace_table_after(text, "## Round 2 — final tree", table("r02_final_"))
text = replace_table_after(text, "### Round 2 — multi-GPU, builder-run mid-round", table("r02_mid_"))
open(path, "w").write(text)
print(text[:600])
