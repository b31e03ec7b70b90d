#!/usr/bin/env python
"""Regenerates tests/golden/reference_kats.json from the reference's own Go tests.

    python tools/gen_kats.py [--check]        (needs /root/reference; --check: compare with the committed file, exit 1 on a difference)

Every number, vector, label map and expected score in the golden file is PARSED out of the literal tables of
/root/reference/{xcorr,muse_batch,muse,group,labels,series,example}_test.go; what is written by hand here is only the
annotation the Go files cannot carry (source line ranges, tolerances the test helpers hard-code, the two tie notes).
tests/test_oracle.py runs the --check when /root/reference exists, so a transcription slip cannot hide in the golden file.
"""
import json
import os
import re
import sys

REF = os.environ.get("MUSE_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "reference_kats.json")


def src(name):
    with open(os.path.join(REF, name)) as f:
        return f.read()


def func_body(text, name):
    """Body of `func name(` by brace matching (strings in these files hold no braces)."""
    i = text.index("func " + name + "(")
    j = text.index("{", text.index(")", i))
    depth, k = 0, j
    while True:
        c = text[k]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return text[j + 1:k]
        k += 1


def line_range(text, name):
    i = text.index("func " + name + "(")
    body = func_body(text, name)
    a = text.count("\n", 0, i) + 1
    b = a + text[i:text.index(body, i) + len(body) + 1].count("\n")      # the line of the closing brace
    return a, b


def num(tok):
    tok = tok.strip()
    return float(tok) if ("." in tok or "e" in tok.lower()) else int(tok)


def floats(s):
    s = s.strip()
    return [num(t) for t in s.split(",") if t.strip()] if s else []


FLOATS = r"\[\]float64\{([^{}]*)\}"
LABELS = r"NewLabels\(\s*LabelMap\{([^{}]*)\}\s*\)"


def label_map(s):
    return {m.group(1): m.group(2) for m in re.finditer(r'"([^"]*)"\s*:\s*"([^"]*)"', s)}


def series_list(body):
    """Every NewSeries([]float64{...}, NewLabels(LabelMap{...})) of a block, in order."""
    out = []
    for m in re.finditer(r"NewSeries\(\s*" + FLOATS + r"\s*,\s*" + LABELS, body):
        out.append({"y": floats(m.group(1)), "labels": label_map(m.group(2))})
    return out


def score_list(block):
    out = []
    for m in re.finditer(r"Score\{Labels:\s*" + LABELS + r",\s*Lag:\s*(-?\d+),\s*PercentScore:\s*(-?[\d.]+)\}", block):
        out.append({"labels": label_map(m.group(1)), "lag": int(m.group(2)), "score": num(m.group(3))})
    return out


def block_after(body, marker):
    """The {...} literal that follows `marker` in body."""
    i = body.index(marker)
    j = body.index("{", i + len(marker) - 1) if not marker.endswith("{") else i + len(marker) - 1
    depth, k = 0, j
    while True:
        if body[k] == "{":
            depth += 1
        elif body[k] == "}":
            depth -= 1
            if depth == 0:
                return body[j:k + 1]
        k += 1


def table_entries(block):
    """Top-level {...} entries of a table literal `{ {..}, {..}, }`."""
    inner, out, depth, start = block[1:-1], [], 0, None
    for k, c in enumerate(inner):
        if c == "{":
            if depth == 0:
                start = k
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                out.append(inner[start:k + 1])
    return out


def sign_of(entry):
    if "isPositive()" in entry:
        return 1
    if "isNegative()" in entry:
        return -1
    assert "x == 0" in entry, entry
    return 0


def results_of(body):
    m = re.search(r"NewResults\((-?\d+),\s*(-?\d+),\s*(-?[\d.]+),\s*SignFilter_(\w+)\)", body)
    sf = {"ANY": 0, "POS": 1, "NEG": -1}[m.group(4)]
    return {"max_lag": int(m.group(1)), "top_n": int(m.group(2)), "threshold": num(m.group(3)), "sign_filter": sf}


def build():
    k = {}
    k["_provenance"] = ("Known-answer vectors of the reference's own Go tests (paths relative to /root/reference), written by tools/gen_kats.py, "
                        "which parses every number, vector and label out of the *_test.go literals (only source ranges, tolerances and the "
                        "tie notes are annotation).  The reference cannot be executed in this image (no Go toolchain), so these expectations "
                        "are what pins the oracle; tests/test_oracle.py re-runs the generator against the committed file when /root/reference "
                        "is present.")
    xt = src("xcorr_test.go")

    def rng(text, fn, name):
        a, b = line_range(text, name)
        return "%s:%d-%d" % (fn, a, b)

    body = func_body(xt, "TestNextPowOf2")
    k["next_pow_of2"] = {"source": rng(xt, "xcorr_test.go", "TestNextPowOf2"),
                         "cases": [[num(m.group(1)), int(m.group(2))] for m in re.finditer(r"\{(-?[\d.]+),\s*(-?\d+)\}", block_after(body, "}{"))]}
    body = func_body(xt, "TestZNormalize")
    k["z_normalize"] = {"source": rng(xt, "xcorr_test.go", "TestZNormalize"), "tol": float(re.search(r"> (1e-\d+)", body).group(1)),
                        "property": "sum(z*z) == len-1",
                        "cases": [floats(m.group(1)) for m in re.finditer(FLOATS, block_after(body, "}{"))]}
    body = func_body(xt, "TestZeroPad")
    cases = []
    for e in table_entries(block_after(body, "}{")):
        m = re.match(r"\{\s*" + FLOATS + r",\s*(\d+),\s*" + FLOATS + r"\s*\}", e)
        cases.append({"x": floats(m.group(1)), "n": int(m.group(2)), "expected": floats(m.group(3))})
    k["zero_pad"] = {"source": rng(xt, "xcorr_test.go", "TestZeroPad"), "cases": cases}

    def xcorr_cases(name, with_norm):
        body = func_body(xt, name)
        out = []
        for e in table_entries(block_after(body, "}{")):
            vecs = [floats(m.group(1)) for m in re.finditer(FLOATS, e)]
            c = {"x": vecs[0], "y": vecs[1]}
            if with_norm:
                c["normalize"] = re.search(r"\b(true|false)\b", e).group(1) == "true"
            c["cc"] = vecs[2] if len(vecs) > 2 else None
            assert (len(vecs) > 2) != bool(re.search(r"\bnil\b", e))
            c["lag"] = int(re.search(r",\s*(-?\d+),\s*(?:isPositive|isNegative|func)", e).group(1))
            c["sign"] = sign_of(e)
            out.append(c)
        return out

    tol = float(re.search(r"> (1e-\d+)", func_body(src("xcorr.go"), "prettyClose")).group(1))
    k["x_corr"] = {"source": rng(xt, "xcorr_test.go", "TestXCorr"), "tol": tol, "cases": xcorr_cases("TestXCorr", True)}
    k["x_corr_with_x"] = {"source": rng(xt, "xcorr_test.go", "TestXCorrWithX"), "tol": tol,
                          "note": "n = len(x) = 5 (not a power of two): ref prepared as znorm(x)/(len-1), zero-padded to n, rfft",
                          "cases": xcorr_cases("TestXCorrWithX", False)}

    bt = src("muse_batch_test.go")
    mt = src("muse_test.go")
    score_tol = float(re.search(r"> (1e-\d+)", func_body(mt, "compareScores")).group(1))

    def batch_case(name, tie=None):
        body = func_body(bt, name)
        ss = series_list(body)
        c = {"source": rng(bt, "muse_batch_test.go", name), "score_tol": score_tol, "ref": ss[0], "comp": ss[1:],
             "results": results_of(body), "concurrency": int(re.search(r"SignFilter_\w+\),\s*(\d+)\)", body).group(1)),
             "group_by": re.findall(r'"([^"]*)"', re.search(r"\.Run\(\[\]string\{([^}]*)\}\)", body).group(1)),
             "expected": score_list(block_after(body, "expectedScores := Scores{"))}
        return c

    c = batch_case("TestBatchRunSimple")
    for e in c["expected"]:
        if e["labels"] == {"graph": "evenLowerShiftedAhead"}:
            e["tie_lags"] = [-3, -2]
    c["tie_note"] = ("evenLowerShiftedAhead: cc[-3] == cc[-2] exactly in exact arithmetic (SURVEY F4); TestRunSimple expects -3 and "
                     "TestRunSimpleSignFilter -2 for the same series, so either lag is the reference's answer")
    k["batch_run_simple"] = c
    k["batch_run_multi_dimensional"] = batch_case("TestBatchRunMultiDimensional")
    body = func_body(bt, "TestBatchRunWithLargerGroup")
    ss = series_list(body)
    assert "err == nil" in body or "err != nil" in body
    k["batch_run_with_larger_group"] = {"source": rng(bt, "muse_batch_test.go", "TestBatchRunWithLargerGroup"), "ref": ss[0], "comp": ss[1:],
                                        "expect_error": "length mismatch"}

    body = func_body(mt, "TestRunSimple")
    ss = series_list(body)
    m = re.search(r"NewResults\((-?\d+),\s*(-?\d+),\s*(-?[\d.]+),", body)
    c = {"source": rng(mt, "muse_test.go", "TestRunSimple"), "score_tol": score_tol, "ref": ss[0], "comp": ss[1:],
         "results": {"max_lag": int(m.group(1)), "top_n": int(m.group(2)), "threshold": num(m.group(3))},
         "expected_any": score_list(block_after(body, "expectedScores := Scores{"))}
    body2 = func_body(mt, "TestRunSimpleSignFilter")
    blocks = [m.start() for m in re.finditer(r"Scores\{", body2)]
    lists = [score_list(block_after(body2[i:], "Scores{")) for i in blocks]
    pos = [l for l in lists if l and all(e["score"] >= 0 for e in l)]
    neg = [l for l in lists if l and all(e["score"] < 0 for e in l)]
    c["expected_pos"], c["expected_neg"] = pos[0], neg[0]
    for lst in (c["expected_any"], c["expected_neg"]):
        for e in lst:
            if e["labels"] == {"graph": "evenLowerShiftedAhead"}:
                e["tie_lags"] = [-3, -2]
    c["source_sign_filter"] = rng(mt, "muse_test.go", "TestRunSimpleSignFilter")
    k["muse_run_simple"] = c

    # ---- labels / group / series: string semantics ----
    st = src("series_test.go")
    y = floats(re.search(r"\by = " + FLOATS, st).group(1))
    gt = src("group_test.go")
    body = func_body(gt, "TestGroupAdd")
    cases = [{"labels": label_map(e), "expect_error": bool(re.search(r"\btrue\b", e))} for e in table_entries(block_after(body, "}{"))]
    k["group_add"] = {"source": rng(gt, "group_test.go", "TestGroupAdd"), "y": y, "cases": cases}
    body = func_body(gt, "TestIndexLabelValues")
    labels = [label_map(m.group(1)) for m in re.finditer(LABELS, body)]
    cases = [[re.findall(r'"([^"]*)"', e), int(re.search(r"(\d+)\s*\}$", e.strip()).group(1))]
             for e in table_entries(block_after(body, "expectedNumLabels int\n\t}{"))]
    body2 = func_body(gt, "TestFilterByLabelValues")
    fcases = [[label_map(e), int(re.search(r"(\d+)\s*\}$", e.strip()).group(1))]
              for e in table_entries(block_after(body2, "expectedNumSeries int\n\t}{"))]
    k["index_label_values"] = {"source": rng(gt, "group_test.go", "TestIndexLabelValues"), "y": y, "labels": labels, "cases": cases,
                               "filter_source": rng(gt, "group_test.go", "TestFilterByLabelValues"), "filter_cases": fcases}
    lt = src("labels_test.go")
    body = func_body(lt, "TestLabelsUID")
    cases = []
    for e in table_entries(block_after(body, "}{")):
        lm = re.search(LABELS, e)
        rest = e[lm.end():]
        m = re.search(r"\[\]string\{([^}]*)\}", rest)
        gb = re.findall(r'"([^"]*)"', m.group(1)) if m else None
        tail = rest[m.end():] if m else rest[rest.index("nil") + 3:]
        cases.append({"labels": label_map(lm.group(1)), "group_by": gb, "expected": re.search(r'"([^"]*)"', tail).group(1)})
    k["labels_id"] = {"source": rng(lt, "labels_test.go", "TestLabelsUID"), "cases": cases}
    a1, _ = line_range(st, "TestNewSeries")
    _, b2 = line_range(st, "TestSeriesLabels")
    ucases = []
    for e in table_entries(block_after(func_body(st, "TestSeriesUID"), "}{")):
        lm = re.search(LABELS, e)
        ucases.append({"labels": label_map(lm.group(1)), "expected": re.search(r'"([^"]*)"', e[lm.end():]).group(1)})
    kcases = []
    for e in table_entries(block_after(func_body(st, "TestSeriesLabels"), "}{")):
        lm = re.search(LABELS, e)
        kcases.append({"labels": label_map(lm.group(1)), "expected": re.findall(r'"([^"]*)"', re.search(r"\[\]string\{([^}]*)\}", e[lm.end():]).group(1))})
    k["series"] = {"source": "series_test.go:%d-%d" % (a1, b2), "default_label": re.search(r'DefaultLabel\s*=\s*"([^"]*)"', src("labels.go")).group(1),
                   "uid_cases": ucases, "key_cases": kcases}

    et = src("example_test.go")
    out = et[et.index("// Output:"):]
    sections, cur = {}, None
    for line in out.splitlines():
        line = line.strip().lstrip("/").strip()
        if line.startswith("Output:"):
            line = line[len("Output:"):].strip()
        m = re.match(r"(.*), Lag: (-?\d+), Score: (-?[\d.]+)$", line)
        if m:
            sections[cur].append({"id": m.group(1), "lag": int(m.group(2)), "score": float(m.group(3))})
        elif line and line != "}":
            cur = line
            sections[cur] = []
    a, b = line_range(et, "Example")
    k["example_structure"] = {
        "source": "example_test.go:%d-%d" % (a, b),
        "note": "Output pinned on Go math/rand data that cannot be regenerated here; only the structure is used: 5 series x 480, maxLag 15, "
                "topN 4, threshold 0; the noise-only ErrorRate/host2 series is absent from 'Unique' (its global peak lies outside +-15: "
                "SURVEY F1) while the constant Line(0,0.1) series is present with score 0.000 lag 0 (xcorr.go:165-168).",
        "expected_unique": sections["Unique"], "expected_by_graph": sections["By Graph"], "expected_by_host": sections["By Host"]}
    return k


def main():
    k = build()
    if "--check" in sys.argv:
        with open(OUT) as f:
            have = json.load(f)
        bad = [name for name in sorted(set(k) | set(have)) if k.get(name) != have.get(name)]
        for name in bad:
            print("DIFFERS:", name)
            print("  generated:", json.dumps(k.get(name))[:600])
            print("  committed:", json.dumps(have.get(name))[:600])
        sys.exit(1 if bad else 0)
    with open(OUT, "w") as f:
        json.dump(k, f, indent=1)
        f.write("\n")
    print("wrote", OUT)


if __name__ == "__main__":
    main()
