"""Generates tests/golden/parser_cases.json by running the REFERENCE's own text -> windows functions
(`post_process`, `moment_str_to_list`; utils/utils.py:66-132, 364-415) on a corpus of LLM-style outputs.

utils/utils.py imports wandb (absent offline), so the two functions are taken out of the module source with `ast` and
executed with their only dependencies (`re`, `ast`) -- the reference code itself runs, unmodified; nothing is copied
into the repository.  Run in the build container only (needs /root/reference):  python tests/golden/make_golden_parser.py
"""
import ast
import json
import os
import random
import re

SRC = "/root/reference/utils/utils.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "parser_cases.json")


def load_reference_functions():
    tree = ast.parse(open(SRC).read())
    ns = {"re": re, "ast": ast}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("post_process", "moment_str_to_list"):
            exec(compile(ast.Module([node], []), SRC, "exec"), ns)
    return ns["post_process"], ns["moment_str_to_list"]


def corpus(seed=11, n_random=3000):
    fixed = [
        "[[0, 1], [4, 7]]", "[[0, 1] [4, 7]]", "[[0 1], [4 7]]", "[[1, 0]]", "[[0,, 1],, [4, 7]]", "[[10, 20]]</s> junk",
        "[[10, 20]]\n", "[[10,\n20]]", "", "no windows", "[]", "[[]]", "[[ ]]", "[[-1, -1]]", "[[3]]", "[[1, 2, 3]]",
        "[[1, 2], 3]", "[[1, 2], [3]]", "[[1.5, 2]]", "[[1, 2.5], [3, 4]]", "[['a', 1]]", "[[0, 1]] trailing", "[[0, 1]], [[2, 3]]",
        "[[0, 1], [2, 3]", "[[0, 1]]]", "[[[0, 1]]]", "[[0, 1], [4, 7]] [[8, 9]]", "[[12, 14],  [20, 30]]", "[[12, 14],[20, 30]]",
        "[[12,14],[20,30]]", "[[ 12 , 14 ]]", "[[12 14] [20 30]]", "[[150, 2]]", "[[2, 150], [148, 150], [0, 0]]",
        "[[True, 2]]", "[[None, 2]]", "[[1, 2], 'ab']", "[[1, 2], 'abc']", "[[1, 2], 2.5]", "[[1, 2], (3, 4)]", "[[1, 2], (3, 'x')]",
        "[[1, 2], {1: 2, 3: 4}]", "[[1, 2], {0: 1, 1: 2}]", "[[1, 2], None]", "[[007, 8]]", "[[1, 2], [3, 4], [5, 6], [7, 8], [9, 10]]",
        "[[1, 2]]\r\n[[3, 4]]", " [[1, 2]]", "[[1, 2]] ", "[[1,2],\t[3,4]]", "[[1 2 3]]", "[[1  2]]", "[[-5, 3]]", "[[3, -5]]",
        "[[1, 2], [4, 3], [6, 5]]", "[[1e2, 3]]", "[[1, 2] , [3, 4]]", "[[1, 2];[3, 4]]", "Relevant windows: [[1, 2]]",
        "[[1, 2], 2.5, [3, 4]]", "[[1, 2], 'ab', [3, 4]]", "[[1, 2], 'abc', [3, 4]]", "[[1, 2], (3, 4), [5, 6]]", "[[1, 2], (3, 'x'), [5, 6]]",
        "[[1, 2], {1: 2, 3: 4}, [5, 6]]", "[[1, 2], {0: 1, 1: 2}, [5, 6]]", "[[1, 2], None, [3, 4]]", "[[1, 2], 7, [3, 4]]",
        "[[1, 2], True, [3, 4]]", "[[1, 2], [3, [4, 5]]]", "[[1, 2], [[3, 4], 5]]", "[[1, 2], b'xy', [3, 4]]", "[[1, 2], {1, 2}, [3, 4]]",
        "[[1, 2]]</s>[[3, 4]]", "[[0x10, 2]]", "[[1_0, 2]]", "[[1, 2], [3, 4],]", "[[1, 2],, ]", "[[1, 2], []]", "[[99999999999, 1]]",
    ]
    rng = random.Random(seed)
    out = list(fixed)
    for _ in range(n_random):
        k = rng.randint(0, 6)
        wins = []
        for _ in range(k):
            a, b = rng.randint(0, 150), rng.randint(0, 150)
            style = rng.random()
            if style < 0.55:
                w = f"[{a}, {b}]"
            elif style < 0.65:
                w = f"[{a} {b}]"
            elif style < 0.72:
                w = f"[{a},, {b}]"
            elif style < 0.78:
                w = f"[{a}, {b}, {rng.randint(0, 9)}]"
            elif style < 0.83:
                w = f"[{a}]"
            elif style < 0.88:
                w = f"[{a}.{rng.randint(0, 9)}, {b}]"
            elif style < 0.92:
                w = f"[{a},{b}]"
            elif style < 0.96:
                w = f"[-{a}, {b}]"
            else:
                w = rng.choice(["[]", "[a, b]", "3", "'x'", "[1, [2, 3]]", "(1, 2)"])
            wins.append(w)
        sep = rng.choice([", ", ", ", ", ", " ", ",", ",, ", " , ", "  "])
        s = "[" + sep.join(wins) + "]"
        r = rng.random()
        if r < 0.08:
            s = s + "</s>" + rng.choice(["", " extra", "[[1, 2]]"])
        elif r < 0.12:
            s = s[:-1]
        elif r < 0.16:
            s = rng.choice(["The windows are ", " ", "\n"]) + s
        elif r < 0.20:
            s = s.replace(" ", "\n", 1)
        elif r < 0.23:
            s = s + rng.choice(["]", " ", ".", "\n"])
        out.append(s)
    return out


def main():
    post_process, moment_str_to_list = load_reference_functions()
    cases = []
    for s in corpus():
        pp = post_process(s)
        rec = {"raw": s, "post": pp}
        for key, arg in (("windows", pp), ("windows_raw", s)):   # evaluate.py:48 composes them; also the bare parser
            try:
                rec[key] = json.loads(json.dumps(moment_str_to_list(arg)))   # tuples -> lists, as the jsonl writer does
            except Exception as e:  # the reference lets these propagate (Appendix A)
                rec[key] = {"raises": type(e).__name__}
        cases.append(rec)
    json.dump(cases, open(OUT, "w"), indent=0)
    n_exc = sum(1 for c in cases if isinstance(c["windows"], dict))
    print(f"wrote {len(cases)} cases to {OUT} ({n_exc} raise in the reference)")


if __name__ == "__main__":
    main()
