"""Program blobs (regengo_b200/csrc/blob.hpp layout) written STRAIGHT from the programs mined out of the reference's
generated Go files (tests/golden/generated_goldens.json, made by mine_goldens.py) -- no product code involved.

The oracle (oracle/rgx_oracle.c) consumes these blobs in tests/test_oracle_goldens.py, which pins it against the
reference's own curated inputs independently of the product's front-end: the instruction listing, class byte sets,
Thompson masks and TDFA tables below are what regengo itself emitted.

One blob carries one instruction array for both methods.  The goldens list the goto-machine per method (MatchBytes
has `goto` where FindBytes has a capture); the find listing is used when there is one, and an Alt whose match
listing pushes what the find listing takes is a greedy loop (compiler.go:566-585 reorders those for MatchBytes).
Patterns whose methods are Thompson + TDFA have no instruction listing at all: their blob holds placeholder
instructions (the two engines never read them) with the char-state flag the Thompson loop looks at.
"""
import json
import os
import struct

HERE = os.path.dirname(os.path.abspath(__file__))
MAGIC, VERSION, HDR = 0x42584752, 1, 32
(H_MAGIC, H_VERSION, H_WORDS, H_NINST, H_START, H_NUMCAP, H_FLAGS, H_PREFIX, H_MATCH_ENGINE, H_FIND_ENGINE, H_MINLEN, H_MAXLEN,
 H_LEFTOVER, H_MINBUF, H_OFF_INST, H_OFF_CLASS, H_OFF_RANGES, H_NRANGE_PAIRS, H_OFF_THOMPSON, H_OFF_TDFA, H_OFF_NAMES, H_NAMES_WORDS) = range(22)
F_ANCHORED, F_NEEDS_BT, F_HAS_PREFIX, F_MATCH_MEMO, F_FIND_MEMO, F_PER_CAPTURE, F_HAS_CAPTURES = 1, 2, 4, 8, 16, 32, 64
OP = {"alt": 0, "cap": 2, "empty": 3, "match": 4, "fail": 5, "goto": 6, "byteclass": 7, "any": 9, "anynotnl": 10}
IF_ALT_CKPT, IF_GREEDY_LOOP, IF_CHAR_STATE = 1, 2, 8


def load_goldens():
    with open(os.path.join(HERE, "generated_goldens.json")) as fh:
        return json.load(fh)


def blob_from_golden(e):
    """-> bytes (little-endian u32 words)."""
    gm, gf = e["match"], e["find"]
    listing = gf if gf and gf["kind"] == "bt" else gm if gm["kind"] == "bt" else None
    n = len(listing["inst"]) if listing else gm["n_inst"]
    w = [0] * HDR
    w[H_MAGIC], w[H_VERSION] = MAGIC, VERSION
    w[H_NINST] = n
    w[H_START] = listing["start"] if listing else 0
    w[H_NUMCAP] = (gf["num_cap"] if gf and gf["kind"] == "bt" else gf["num_tags"] if gf else 2) or 2
    flags = 0
    anchored = (not gm["retry"]) if gm["kind"] == "bt" else gm["anchored"]
    if anchored:
        flags |= F_ANCHORED
    if listing and listing["uses_stack"]:
        flags |= F_NEEDS_BT
    prefix = gm.get("prefix") if gm["kind"] == "bt" else (gf.get("prefix") if gf and gf["kind"] == "tdfa" else None)
    if prefix is not None:
        flags |= F_HAS_PREFIX
    if gm["kind"] == "bt" and gm["memo"]:
        flags |= F_MATCH_MEMO
    if gf and gf["kind"] == "bt" and (gf["memo"] or e.get("findall_memo")):
        flags |= F_FIND_MEMO
    if gf and gf["kind"] == "bt" and gf["per_capture"]:
        flags |= F_PER_CAPTURE
    if gf:
        flags |= F_HAS_CAPTURES
    w[H_FLAGS] = flags
    w[H_PREFIX] = prefix or 0
    w[H_MATCH_ENGINE] = 0 if gm["kind"] == "bt" else 1
    w[H_FIND_ENGINE] = 0 if not gf else 1 if gf["kind"] == "bt" else 2
    w[H_MINLEN], w[H_MAXLEN] = (e["min_match_len"] or 0) & 0xFFFFFFFF, (e["max_match_len"] if e["max_match_len"] is not None else -1) & 0xFFFFFFFF
    w[H_LEFTOVER], w[H_MINBUF] = e["default_max_leftover"] or 0, e["min_buffer"] or 0

    # instructions + class bitmaps
    inst, cls = [], []
    char_states = set(int(s) for s in gm["cond"]) if gm["kind"] == "thompson" else set()
    for i in range(n):
        bits = [0] * 8
        if listing is None:
            inst += [OP["fail"] | ((IF_CHAR_STATE if i in char_states else 0) << 8), 0, 0, 0]
            cls += bits
            continue
        gi = listing["inst"][i]
        k = gi["kind"]
        fl = IF_CHAR_STATE if i in char_states else 0
        out, arg = gi.get("out", 0), gi.get("arg", 0)
        if k == "alt":
            out, arg = gi["taken"], gi["pushed"]
            if listing is gf:
                if gi.get("ckpt"):
                    fl |= IF_ALT_CKPT
                if gm["kind"] == "bt" and gm["inst"][i]["kind"] == "alt" and gm["inst"][i]["pushed"] == out and gm["inst"][i]["taken"] == arg and out != arg:
                    fl |= IF_GREEDY_LOOP
        elif k == "byteclass":
            for c in gi["set"]:
                bits[c >> 5] |= 1 << (c & 31)
        inst += [OP[k] | (fl << 8), out, arg, 0]
        cls += bits
    w[H_OFF_INST] = len(w)
    w += inst
    w[H_OFF_CLASS] = len(w)
    w += cls
    w[H_OFF_RANGES] = len(w)
    w += [0, 0] * n
    w[H_NRANGE_PAIRS] = 0
    # Thompson masks
    w[H_OFF_THOMPSON] = len(w)

    def put64(v):
        v = int(v)
        return [v & 0xFFFFFFFF, (v >> 32) & 0xFFFFFFFF]
    if gm["kind"] == "thompson":
        w += put64(gm["start_closure"]) + put64(gm["accept_mask"])
        for i in range(n):
            w += put64(gm["eps_after"].get(str(i), 0))
        for i in range(n):
            bits = [0] * 8
            for c in gm["cond"].get(str(i), []):
                bits[c >> 5] |= 1 << (c & 31)
            w += bits
    else:
        w += [0] * (4 + 2 * n + 8 * n)
    # TDFA tables
    if gf and gf["kind"] == "tdfa":
        ns = len(gf["transitions"])
        base = len(w)
        w[H_OFF_TDFA] = base
        hdr = [0] * 12
        hdr[0], hdr[1], hdr[2], hdr[3] = ns, gf["num_tags"], gf["start_begin"], gf["start_any"]
        hdr[4], hdr[5] = len(gf["init_tags_begin"]), len(gf["init_tags_any"])
        acts_off, acts = [], []
        for s in range(ns):
            for c in range(128):
                acts_off.append(len(acts) // 2)
                cnt = gf["tagActionCount"][s][c] if gf["tagActionCount"] else 0
                for a in range(cnt):
                    acts += [gf["tagActionTags"][s][c][a], gf["tagActionOffsets"][s][c][a]]
        acts_off.append(len(acts) // 2)
        acc_off, acc = [], []
        for s in range(ns):
            acc_off.append(len(acc) // 2)
            cnt = gf["acceptActionCount"][s] if gf["acceptActionCount"] else 0
            for a in range(cnt):
                acc += [gf["acceptActionTags"][s][a], gf["acceptActionOffsets"][s][a]]
        acc_off.append(len(acc) // 2)
        hdr[6], hdr[7] = len(acts) // 2, len(acc) // 2
        hdr[8] = max([gf["tagActionCount"][s][c] for s in range(ns) for c in range(128)], default=0) if gf["tagActionCount"] else 0
        hdr[9] = max(gf["acceptActionCount"], default=0) if gf["acceptActionCount"] else 0
        w += hdr + list(gf["init_tags_begin"]) + list(gf["init_tags_any"])
        w += [x & 0xFFFFFFFF for row in gf["transitions"] for x in row]
        w += [int(x) for x in gf["acceptStates"]] + [int(x) for x in gf["acceptStatesEOT"]]
        w += acts_off + acts + acc_off + acc
    # names (unnamed: the oracle does not read them)
    w[H_OFF_NAMES] = len(w)
    w[H_NAMES_WORDS] = 0
    w[H_WORDS] = len(w)
    return struct.pack("<%dI" % len(w), *w)
