"""Generates the committed golden fixtures from the reference checkout (run in the authoring
container only -- /root/reference does not exist on the GPU box):

  datalist_{2016,2012,2008}.npz   the named data list final_{year}.R hands to Stan
                                  (built by us-potus-model_b200/datalist.py from /root/reference/data)
  readme_tables.json              the reference's published election-day tables
                                  (README.md:83-136 2008, :179-232 2012, :279-332 2016): state, mean, low, high, prob, se
  known_answers.json              lp / |grad| known-answer values re-derived
                                  here with the fp64 oracle, plus data-list anchors

usage: python tests/golden/make_fixtures.py [/root/reference]
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import potus_pkg  # noqa: E402

pkg = potus_pkg.load()
import potus_oracle as po  # noqa: E402


def parse_readme_tables(path):
    lines = open(path).read().split("\n")
    tables, cur = [], None
    for ln in lines:
        if re.match(r"^\|\s*state\s*\|\s*mean\s*\|\s*low\s*\|\s*high\s*\|\s*prob\s*\|\s*se\s*\|", ln):
            cur = []
            tables.append(cur)
            continue
        if cur is not None:
            if ln.startswith("|"):
                cells = [c.strip() for c in ln.strip().strip("|").split("|")]
                if set(cells[0]) <= set(":-"):
                    continue
                cur.append(dict(state=cells[0], mean=float(cells[1]), low=float(cells[2]), high=float(cells[3]),
                                prob=float(cells[4]), se=float(cells[5])))
            else:
                cur = None
    assert len(tables) == 3, len(tables)
    return {"2008": tables[0], "2012": tables[1], "2016": tables[2]}


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    kat = {}
    for year in (2016, 2012, 2008):
        d = pkg.build_datalist(year, os.path.join(ref, "data"))
        pkg.save_npz(os.path.join(HERE, f"datalist_{year}.npz"), d)
        _, D = po.block_layout(d)
        th0 = np.zeros(D)
        th1 = 0.1 * np.sin(1 + 0.37 * np.arange(D))
        lp0, g0 = po.logp_grad_closed(th0, d)
        lp1, g1 = po.logp_grad_closed(th1, d)
        kat[str(year)] = dict(D=D, lp_zero=lp0, gnorm_zero=float(np.linalg.norm(g0)), lp_sin=lp1,
                              gnorm_sin=float(np.linalg.norm(g1)), g_zero_first=float(g0[0]), g_zero_last=float(g0[-1]),
                              N_state_polls=int(d["N_state_polls"]), N_national_polls=int(d["N_national_polls"]),
                              T=int(d["T"]), P=int(d["P"]),
                              sum_y_state=int(d["n_democrat_state"].sum()), sum_n_state=int(d["n_two_share_state"].sum()),
                              sum_y_nat=int(d["n_democrat_national"].sum()), sum_n_nat=int(d["n_two_share_national"].sum()))
    json.dump(kat, open(os.path.join(HERE, "known_answers.json"), "w"), indent=1)
    json.dump(parse_readme_tables(os.path.join(ref, "README.md")), open(os.path.join(HERE, "readme_tables.json"), "w"), indent=0)
    print("wrote fixtures to", HERE)


if __name__ == "__main__":
    main()
