"""Regenerates tests/golden/reference_vignettes.json from the reference's OWN rendered vignettes.

The reference (an R package) cannot run in this image (no R), so the only exact outputs it gives
for the hot path are the numbers printed in its pre-rendered pkgdown articles.  This script reads
them out of /root/reference/docs/articles/*.html.  It is only needed to regenerate the fixture; the
tests read the committed JSON and never touch /root/reference.

    python tests/golden/extract_goldens.py [/root/reference]
"""
import html
import json
import os
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"


def text_of(name):
    raw = open(os.path.join(ref, "docs", "articles", name), encoding="utf-8").read()
    return html.unescape(re.sub(r"<[^>]*>", "", raw))


ex = text_of("example.html")
rates = [float(v) for v in re.findall(r"acceptance rate: ([0-9.]+)%", ex)]
assert len(rates) == 20, len(rates)
head = []
for m in re.finditer(r"^#>\s+\d+\s+1\s+burnin\s+(\d+)\s+([-0-9.]+)\s+([-0-9.]+)\s+([-0-9.]+)\s+([-0-9.]+)\s*$", ex, re.M):
    head.append([int(m.group(1))] + [float(m.group(k)) for k in range(2, 6)])
head = head[:6]
assert len(head) == 6, head
rhat = re.search(r"mcmc\$diagnostics\$rhat\s*#>\s+mu\s+sigma\s*#>\s+([0-9.]+)\s+([0-9.]+)", ex)
ess = re.search(r"mcmc\$diagnostics\$ess\s*#>\s+mu\s+sigma\s*#>\s+([0-9.]+)\s+([0-9.]+)", ex)

mc = text_of("metropolis_coupling.html")
mc_rates = [float(v) for v in re.findall(r"acceptance rate: ([0-9.]+)%", mc)]
dw = text_of("checks_double_well.html")
dw_rates = [float(v) for v in re.findall(r"acceptance rate: ([0-9.]+)%", dw)]

gm = text_of("getting_model_fits.html")
gm_rates = [float(v) for v in re.findall(r"acceptance rate: ([0-9.]+)%", gm)]
pl = text_of("parallel.html")
pl_rates = [float(v) for v in re.findall(r"acceptance rate: ([0-9.]+)%", pl)]

out = {
    "source": "mrc-ide/drjacoby docs/articles (pkgdown build 2024-06-26), package version 1.5.4",
    "example": {
        "where": "docs/articles/example.html:275-330,359,384,520-555; setup vignettes/example.Rmd:19-51,120-129,204-213",
        "setup": "set.seed(1); x = rnorm(10, 3, 2); mu in [-10,10], sigma in [0,Inf); no init; 5 chains; burnin 1e3; samples 1e3; run 1 (R closures) then run 2 (C++), one RNG stream",
        "head_rows_chain1": {"columns": ["iteration", "mu", "sigma", "logprior", "loglikelihood"], "rows": head},
        "acceptance_pct_run1": [[rates[2 * i], rates[2 * i + 1]] for i in range(5)],
        "acceptance_pct_run2": [[rates[10 + 2 * i], rates[10 + 2 * i + 1]] for i in range(5)],
        "rhat": [float(rhat.group(1)), float(rhat.group(2))],
        "ess": [float(ess.group(1)), float(ess.group(2))],
    },
    "metropolis_coupling": {
        "where": "docs/articles/metropolis_coupling.html:284,287; setup vignettes/metropolis_coupling.Rmd:21,52-61,70-107",
        "setup": "set.seed(1); x = rnorm(100, 10, 1); init (5,5,0); nine runs 1 rung 1e3+1e4; then rungs=20, 1e3+1e4",
        "acceptance_pct_all_runs": mc_rates,
    },
    "getting_model_fits": {
        "where": "docs/articles/getting_model_fits.html:247-266; setup vignettes/getting_model_fits.Rmd:21,61-63,72-116",
        "setup": "set.seed(1); population_data; N_0 in [1,100], r in [1,2], N_max in [50,200], no init; logistic map + dpois (R closures); 3 chains, 1e3+1e4",
        "acceptance_pct_all_runs": gm_rates,
    },
    "parallel": {
        "where": "docs/articles/parallel.html:181-192; setup vignettes/parallel.Rmd:21,36-66",
        "setup": "set.seed(1); x = rnorm(10); mu in [-10,10], init 0; sum dnorm(x, mu, 1); 2 chains, 1e3+1e3",
        "acceptance_pct_all_runs": pl_rates,
    },
    "double_well": {
        "where": "docs/articles/checks_double_well.html:203,206; setup vignettes/checks_double_well.Rmd:18,46-65,91-101",
        "setup": "set.seed(1); mu in [-2,2], gamma fixed 30; run 1: 1 rung 1e3+1e5 (silent); run 2: rungs=11, alpha=2, 1e3+1e5",
        "acceptance_pct_all_runs": dw_rates,
    },
}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vignettes.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
