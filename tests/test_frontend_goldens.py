"""Front-end pin: the programs our restated regexp/syntax + internal/compiler analysis produce must
equal, instruction by instruction and table cell by table cell, what the reference itself baked into
its 24 checked-in generated Go files (mined by tests/golden/mine_goldens.py into
tests/golden/generated_goldens.json).  Covers: Prog layout (op, out/arg, class byte sets), Match-mode
greedy-loop reordering, checkpoint flags, memoisation, prefix byte, Thompson masks + per-state byte
conditions, full TDFA tables (transitions, tag actions, accept sets/actions, start states), stream
constants."""
import pytest

from helpers import compile_json

OPN = {0: "alt", 1: "altmatch", 2: "cap", 3: "empty", 4: "match", 5: "fail", 6: "nop", 7: "rune", 8: "rune1", 9: "any", 10: "anynotnl"}


def bits_to_set(bits, i):
    return [c for c in range(256) if (bits[i * 8 + (c >> 5)] >> (c & 31)) & 1]


def cmp_bt(j, gm, mode, msgs):
    if gm["start"] != j["start"]:
        msgs.append(f"{mode} start {j['start']} != {gm['start']}")
    if len(gm["inst"]) != len(j["inst"]):
        msgs.append(f"{mode} n_inst {len(j['inst'])} != {len(gm['inst'])}")
        return
    memo = j["match_memo"] if mode == "match" else j["find_memo"]
    if gm["memo"] != bool(memo):
        msgs.append(f"{mode} memo {memo} != {gm['memo']}")
    if gm["memo"] and gm["memo_ninst"] is not None and gm["memo_ninst"] != len(j["inst"]):
        msgs.append("memo numInst")
    if gm["uses_stack"] != bool(j["needs_backtracking"]):
        msgs.append("needs_backtracking")
    if gm["retry"] == bool(j["anchored"]):
        msgs.append("anchored/retry")
    if mode == "match":
        pref = j["prefix"] if (j["has_prefix"] and not j["anchored"]) else None
        if gm["prefix"] != pref:
            msgs.append(f"prefix {pref} != {gm['prefix']}")
    else:
        if j["needs_backtracking"] and gm["per_capture"] != bool(j["per_capture_ckpt"]):
            msgs.append(f"per_capture {j['per_capture_ckpt']} != {gm['per_capture']}")
        if gm["num_cap"] != j["num_cap"]:
            msgs.append(f"num_cap {j['num_cap']} != {gm['num_cap']}")
    for i, (gi, ji) in enumerate(zip(gm["inst"], j["inst"])):
        op, k = OPN[ji["op"]], gi["kind"]
        ok = True
        if op in ("fail", "match"):
            ok = k == op
        elif op == "nop":
            ok = k == "goto" and gi["out"] == ji["out"]
        elif op == "cap":
            if mode == "match":
                ok = k == "goto" and gi["out"] == ji["out"]
            else:
                ok = k == "cap" and gi["arg"] == ji["arg"] and gi["out"] == ji["out"] and gi["percap"] == bool(j["per_capture_ckpt"])
        elif op == "rune1":
            r = ji["rune"][0]
            ok = (k == "byteclass" and gi["set"] == [r] and gi["out"] == ji["out"]) if r < 128 else k in ("unicode", "byteclass")
        elif op == "rune":
            if j["unicode_class"][i]:
                ok = k == "unicode"
            else:
                ok = k == "byteclass" and gi["set"] == bits_to_set(j["class_bits"], i) and gi["out"] == ji["out"]
        elif op in ("any", "anynotnl"):
            ok = k == op and gi["out"] == ji["out"]
        elif op == "empty":
            ok = k == "empty" and gi["arg"] == ji["arg"] and gi["out"] == ji["out"]
        elif op == "alt":
            if k != "alt":
                ok = False
            else:
                if mode == "match" and j["greedy_loop"][i]:
                    ok = gi["pushed"] == ji["out"] and gi["taken"] == ji["arg"]
                else:
                    ok = gi["pushed"] == ji["arg"] and gi["taken"] == ji["out"]
                if mode == "find" and not j["per_capture_ckpt"]:
                    ok = ok and gi["flag"] == j["alt_ckpt"][i] and gi["ckpt"] == bool(j["alt_ckpt"][i])
                if mode == "find" and j["per_capture_ckpt"]:
                    ok = ok and gi["flag"] == 0
                ok = ok and gi["memo"] == bool(memo)
        if not ok:
            msgs.append(f"{mode} inst {i}: ours {ji} ({op}) vs golden {gi}")


def compare(e):
    j = compile_json(e["pattern"])
    msgs = []
    for k in ("min_match_len", "max_match_len", "default_max_leftover", "min_buffer"):
        if e[k] is not None and j[k] != e[k]:
            msgs.append(f"{k}: {j[k]} != {e[k]}")
    gm = e["match"]
    if gm["kind"] == "bt":
        if j["match_engine"] != 0:
            msgs.append("match engine: ours thompson, golden bt")
        else:
            cmp_bt(j, gm, "match", msgs)
    else:
        if j["match_engine"] != 1:
            msgs.append("match engine: ours bt, golden thompson")
        else:
            if str(gm["start_closure"]) != j["start_closure"]:
                msgs.append(f"start_closure {j['start_closure']} != {gm['start_closure']}")
            if str(gm["accept_mask"]) != j["accept_mask"]:
                msgs.append("accept_mask")
            if gm["n_inst"] != len(j["inst"]):
                msgs.append("thompson n_inst")
            for s, v in gm["eps_after"].items():
                if j["eps_after"][int(s)] != v:
                    msgs.append(f"eps_after[{s}] {j['eps_after'][int(s)]} != {v}")
            char_states = [i for i, c in enumerate(j["char_state"]) if c]
            if sorted(int(s) for s in gm["cond"]) != char_states:
                msgs.append(f"char states {char_states} != {sorted(gm['cond'])}")
            for s, cs in gm["cond"].items():
                if bits_to_set(j["thompson_cond"], int(s)) != cs:
                    msgs.append(f"thompson cond[{s}]")
            if bool(j["anchored"]) != gm["anchored"]:
                msgs.append("thompson anchored")
    gf = e["find"]
    if gf is None:
        if j["find_engine"] != 0:
            msgs.append("find engine should be none")
    elif gf["kind"] == "bt":
        if j["find_engine"] != 1:
            msgs.append(f"find engine: ours {j['find_engine']}, golden bt")
        else:
            cmp_bt(j, gf, "find", msgs)
            if e.get("findall_memo") is not None and e["findall_memo"] != bool(j["find_memo"]):
                msgs.append("findall memo")
    else:
        if j["find_engine"] != 2:
            msgs.append(f"find engine: ours {j['find_engine']}, golden tdfa")
        else:
            t = j["tdfa"]
            ns = t["num_states"]
            if ns != len(gf["transitions"]):
                msgs.append(f"tdfa states {ns} != {len(gf['transitions'])}")
            else:
                if [x for row in gf["transitions"] for x in row] != t["trans"]:
                    msgs.append("tdfa transitions differ")
                if [int(x) for x in gf["acceptStates"]] != t["accept"]:
                    msgs.append("accept differ")
                if [int(x) for x in gf["acceptStatesEOT"]] != t["accept_eot"]:
                    msgs.append("accept_eot differ")
                for k2 in ("num_tags", "start_begin", "start_any", "init_tags_begin", "init_tags_any"):
                    if gf[k2] != t[k2]:
                        msgs.append(f"tdfa {k2}: {t[k2]} != {gf[k2]}")
                for s in range(ns):
                    for c in range(128):
                        cnt = gf["tagActionCount"][s][c] if gf["tagActionCount"] else 0
                        ga = [[gf["tagActionTags"][s][c][a], gf["tagActionOffsets"][s][c][a]] for a in range(cnt)]
                        if ga != t["actions"].get(str(s * 128 + c), []):
                            msgs.append(f"tdfa action [{s}][{c}]")
                    cnt = gf["acceptActionCount"][s] if gf["acceptActionCount"] else 0
                    ga = [[gf["acceptActionTags"][s][a], gf["acceptActionOffsets"][s][a]] for a in range(cnt)]
                    if ga != t["accept_actions"][s]:
                        msgs.append(f"tdfa accept action [{s}]")
            pref = j["prefix"] if (j["has_prefix"] and not j["anchored"]) else None
            if gf["prefix"] != pref:
                msgs.append(f"tdfa prefix {pref} != {gf['prefix']}")
    return msgs


def test_all_goldens_present(goldens):
    assert len(goldens) == 24
    assert sum(1 for e in goldens if e["find"] and e["find"]["kind"] == "tdfa") == 5


@pytest.mark.parametrize("idx", range(24))
def test_frontend_reproduces_generated_program(goldens, idx):
    e = goldens[idx]
    msgs = compare(e)
    assert not msgs, e["file"] + ": " + "; ".join(msgs[:6])
