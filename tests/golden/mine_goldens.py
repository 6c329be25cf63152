#!/usr/bin/env python3
"""Mine the reference's checked-in GENERATED Go files into JSON fixtures.

Run here (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/mine_goldens.py            -> tests/golden/generated_goldens.json

The reference has no Prog dumps; its generated files (benchmarks/curated/*.go,
tests/integration/streaming/testdata/*_pattern.go, benchmarks/streams/testdata/*.go,
examples/streaming/date_pattern.go) literally contain the program: one `InsN:` block per
syntax.Inst (internal/compiler/instructions.go), Thompson masks (thompson.go:86-154) and
TDFA tables (tdfa.go:547-794).  This script extracts them without interpreting the regex:
byte-class conditions are *evaluated* for c in 0..255 after a textual Go->Python rewrite.
"""
import glob
import json
import os
import re
import sys

REF = "/root/reference"
FILES = sorted(
    glob.glob(f"{REF}/benchmarks/curated/*.go")
    + glob.glob(f"{REF}/tests/integration/streaming/testdata/*_pattern.go")
    + glob.glob(f"{REF}/benchmarks/streams/testdata/*.go")
    + glob.glob(f"{REF}/examples/streaming/*_pattern.go")
)
FILES = [f for f in FILES if not f.endswith("_test.go") and not f.endswith("generate.go")]


def func_body(src, recv, name):
    m = re.search(r"^func \((?:\w+ )?%s\) %s\(.*?\n(.*?)^}\n" % (re.escape(recv), name), src, re.S | re.M)
    return m.group(1) if m else None


def go_cond_to_py(cond):
    c = cond
    c = re.sub(r"\[32\]byte\{([^}]*)\}", lambda m: "[" + m.group(1) + "]", c)
    c = re.sub(r"uint8\((0x[0-9a-fA-F]+)\)", r"\1", c)
    c = c.replace("input[offset]", "c").replace("||", " or ").replace("&&", " and ")
    c = c.replace("/", "//")
    c = re.sub(r"!(?!=)", " not ", c)
    return c


def byte_set(fail_cond):
    py = go_cond_to_py(fail_cond)
    code = compile(py, "<cond>", "eval")
    return [c for c in range(256) if not eval(code, {"c": c})]


def parse_blocks(body):
    """Split a goto-machine body into {n: text} for the InsN blocks."""
    parts = re.split(r"^Ins(\d+):\n", body, flags=re.M)
    blocks = {}
    for i in range(1, len(parts), 2):
        blocks[int(parts[i])] = parts[i + 1]
    return blocks


def classify(text, mode):
    """mode: 'match' | 'find'.  Returns a dict describing the instruction."""
    t = text
    d = {}
    m_push = re.search(r"stack = append\(stack, \[(\d)\]int\{([^}]*)\}\)", t)
    gotos = re.findall(r"goto Ins(\d+)", t)
    if "DecodeRune" in t:
        d["kind"] = "unicode"
        return d
    if m_push and "captures[" in m_push.group(2) and m_push.group(2).strip().endswith("2"):
        # per-capture checkpoint Capture inst: [3]int{captures[k], k, 2}
        k = int(re.search(r"captures\[(\d+)\] = offset", t).group(1))
        d.update(kind="cap", arg=k, out=int(re.search(r"nextInstruction = (\d+)", t).group(1)), percap=True)
        return d
    if m_push:
        fields = [x.strip() for x in m_push.group(2).split(",")]
        d.update(kind="alt", pushed=int(fields[1]), taken=int(gotos[-1]),
                 flag=int(fields[2]) if len(fields) > 2 else 0,
                 ckpt="captureStack = append" in t, memo="visited[word]" in t)
        return d
    m = re.search(r"captures\[(\d+)\] = offset\n\s*nextInstruction = (\d+)", t)
    if m and "captures[1] = offset" not in t.split("nextInstruction")[0][:0] and int(m.group(1)) != 1:
        d.update(kind="cap", arg=int(m.group(1)), out=int(m.group(2)), percap=False)
        return d
    if m and int(m.group(1)) == 1 and "r.Match" not in t and "item.Match" not in t:
        d.update(kind="cap", arg=1, out=int(m.group(2)), percap=False)
        return d
    if "captures[1] = offset" in t or re.search(r"^\s*return true\s*$", t, re.M):
        if "if" not in t.split("return true")[0] if "return true" in t else True:
            d["kind"] = "match"
            return d
    m = re.search(r"if l <= offset( \|\| input\[offset\] == uint8\(0xa\))? \{\n\s*goto TryFallback\n\s*\}\n(.*)", t, re.S)
    if m:
        rest = m.group(2)
        if m.group(1):
            d.update(kind="anynotnl", out=int(gotos[-1]))
            return d
        mc = re.search(r"if (.*) \{\n\s*goto TryFallback", rest)
        if mc is None:
            d.update(kind="any", out=int(gotos[-1]))
            return d
        d.update(kind="byteclass", set=byte_set(mc.group(1)), out=int(gotos[-1]))
        return d
    if "prevIsWord" in t or "offset != 0" in t or "offset != l" in t:
        ops = 0
        if "prevIsWord == currIsWord" in t:
            ops |= 16
        if "prevIsWord != currIsWord" in t:
            ops |= 32
        if re.search(r"if offset != 0 \{", t):
            ops |= 4
        if re.search(r"if offset != l \{", t):
            ops |= 8
        if "offset != 0 && (offset == 0" in t:
            ops |= 1
        if "offset != l && input[offset]" in t:
            ops |= 2
        d.update(kind="empty", arg=ops, out=int(gotos[-1]))
        return d
    body = t.strip()
    if re.fullmatch(r"\{\s*return (false|nil, false)\s*\}", body):
        d["kind"] = "fail"
        return d
    if re.fullmatch(r"\{\s*goto TryFallback\s*\}", body):
        d["kind"] = "fail"
        return d
    m = re.fullmatch(r"\{\s*goto Ins(\d+)\s*\}", body)
    if m:
        d.update(kind="goto", out=int(m.group(1)))  # Nop, or Capture skipped in Match mode
        return d
    raise ValueError("unclassified block:\n" + t[:400])


def mine_bt(body, mode):
    out = {}
    out["start"] = int(re.search(r"nextInstruction := (\d+)", body).group(1))
    head = body.split("StepSelect:")[0]
    out["memo"] = "visited" in head
    m = re.search(r"visitedSize := (\d+) \*", head)
    out["memo_ninst"] = int(m.group(1)) if m else None
    m = re.search(r"bytes\.IndexByte\(input, uint8\((0x[0-9a-f]+)\)\)", head)
    out["prefix"] = int(m.group(1), 16) if m else None
    fb = head.split("TryFallback:")[1]
    out["retry"] = "l > offset" in fb or "searchStart++" in fb
    out["per_capture"] = "last[2] == 2" in fb
    out["uses_stack"] = "len(stack)" in fb
    m = re.search(r"var captures \[(\d+)\]int", head)
    out["num_cap"] = int(m.group(1)) if m else None
    blocks = parse_blocks(body.split("StepSelect:")[1])
    out["inst"] = [classify(blocks[i], mode) for i in range(len(blocks))]
    return out


def mine_thompson(body):
    out = {}
    out["start_closure"] = str(int(re.search(r"startClosure := uint64\(uint64\((0x[0-9a-f]+)\)\)", body).group(1), 16))
    out["accept_mask"] = str(int(re.search(r"acceptMask := uint64\(uint64\((0x[0-9a-f]+)\)\)", body).group(1), 16))
    m = re.search(r"epsilonClosures := \[(\d+)\]uint64\{(.*?)\}\n", body)
    out["n_inst"] = int(m.group(1))
    eps = {}
    for k, v in re.findall(r"(\d+): uint64\(uint64\((0x[0-9a-f]+)\)\)", m.group(2)):
        eps[k] = str(int(v, 16))
    out["eps_after"] = eps
    out["anchored"] = "Anchored pattern" in body
    # per-state byte conditions
    conds = {}
    for bit, cond in re.findall(r"if current&uint64\(uint64\((0x[0-9a-f]+)\)\) != 0 && (.*) \{\n", body):
        state = int(bit, 16).bit_length() - 1
        py = go_cond_to_py(cond)
        code = compile(py, "<cond>", "eval")
        conds[str(state)] = [c for c in range(256) if eval(code, {"c": c, "true": True, "false": False})]
    out["cond"] = conds
    return out


def go_array(src, name):
    m = re.search(r"^var %s = (.*)$" % re.escape(name), src, re.M)
    if not m:
        return None
    txt = re.sub(r"\s*//.*$", "", m.group(1))
    txt = re.sub(r"(\[\d+\])+(int|bool)\{", "[", txt)
    txt = txt.replace("{", "[").replace("}", "]").replace("true", "True").replace("false", "False")
    return eval(txt)


def mine_tdfa(src, name, body):
    out = {}
    for base in ["transitions", "tagActionCount", "tagActionTags", "tagActionOffsets", "acceptStates",
                 "acceptStatesEOT", "acceptActionCount", "acceptActionTags", "acceptActionOffsets"]:
        out[base] = go_array(src, base + name)
    out["num_tags"] = int(re.search(r"var tags \[(\d+)\]int", body).group(1))
    m = re.search(r"if start == 0 \{\n(.*?)\} else \{\n(.*?)\}\n", body, re.S)
    out["start_begin"] = int(re.search(r"state = (\d+)", m.group(1)).group(1))
    out["start_any"] = int(re.search(r"state = (\d+)", m.group(2)).group(1))
    out["init_tags_begin"] = [int(x) for x in re.findall(r"tags\[(\d+)\] = start", m.group(1))]
    out["init_tags_any"] = [int(x) for x in re.findall(r"tags\[(\d+)\] = start", m.group(2))]
    m = re.search(r"bytes\.IndexByte\(input\[start:\], uint8\((0x[0-9a-f]+)\)\)", body)
    out["prefix"] = int(m.group(1), 16) if m else None
    return out


def mine_file(path):
    src = open(path, encoding="utf-8").read()
    pat = re.search(r"^// Code generated by regengo for pattern: (.*)$", src, re.M).group(1)
    name = re.search(r"^type (\w+) struct\{\}", src, re.M).group(1)
    g = {"file": os.path.relpath(path, REF), "name": name, "pattern": pat}
    g["min_match_len"] = int(re.search(r"const %sMinMatchLen = (-?\d+)" % name, src).group(1))
    g["max_match_len"] = int(re.search(r"const %sMaxMatchLen = (-?\d+)" % name, src).group(1))
    m = re.search(r"func \(%s\) DefaultMaxLeftover\(\) int \{\n\s*return (\d+)" % name, src)
    g["default_max_leftover"] = int(m.group(1)) if m else None
    m = re.search(r"cfg\.Validate\((\d+)\)", src)
    g["min_buffer"] = int(m.group(1)) if m else None
    mb = func_body(src, name, "MatchBytes")
    if "Thompson NFA" in mb:
        g["match"] = dict(kind="thompson", **mine_thompson(mb))
    else:
        g["match"] = dict(kind="bt", **mine_bt(mb, "match"))
    fi = func_body(src, name, "findBytesInternal")
    fr = func_body(src, name, "FindBytesReuse")
    if fi is not None:
        g["find"] = dict(kind="tdfa", **mine_tdfa(src, name, fi))
    elif fr is not None:
        g["find"] = dict(kind="bt", **mine_bt(fr, "find"))
        fa = func_body(src, name, "FindAllBytesAppend")
        g["findall_memo"] = "visited" in fa.split("StepSelect:")[0]
    else:
        g["find"] = None
    return g


def main():
    out = []
    for f in FILES:
        out.append(mine_file(f))
        print("mined", out[-1]["file"], out[-1]["match"]["kind"], (out[-1]["find"] or {}).get("kind"), file=sys.stderr)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "generated_goldens.json")
    with open(dst, "w") as fh:
        json.dump(out, fh, separators=(",", ":"))
    print("wrote", dst, len(out), "files", file=sys.stderr)


if __name__ == "__main__":
    main()
