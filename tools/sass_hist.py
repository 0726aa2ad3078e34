"""Static opcode histogram of one kernel of a cubin / shared library (no GPU needed).

usage: python tools/sass_hist.py <file.cubin|.so> <substring of the mangled or demangled kernel name> [--top N]

The FDCT kernel is straight-line code, so its static counts are (nearly) its executed counts; for
the kernels with loops the histogram only says what the loop bodies are made of.
"""
import collections
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    name, body = None, []
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
        elif re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            body.append(line)
    if name:
        yield name, body


def main():
    path, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    for name, body in kernels(path):
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        if pat not in name and pat not in dem:
            continue
        ops = collections.Counter()
        for line in body:
            text = re.sub(r"^\s+/\*[0-9a-f]+\*/\s+", "", line)
            text = re.sub(r"^@!?U?P\w+\s+", "", text)
            ops[text.split()[0].rstrip(";").split(".")[0]] += 1
        print("%s\n  %d instructions" % (dem[:140], sum(ops.values())))
        print("  " + "  ".join("%s %d" % kv for kv in ops.most_common(top)))


if __name__ == "__main__":
    main()
