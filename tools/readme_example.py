"""Runs the usage example of README.md (kept in sync by hand) and prints the diagnostics."""
import os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "README.md")).read()
code = re.search(r"## Using it like the reference\s+```python\n(.*?)```", text, flags=re.S).group(1)
ns = {}
exec(code, ns)
m = ns["mcmc"]
print(m.output.head())
print("rhat", m.diagnostics["rhat"].to_dict(), "ess", m.diagnostics["ess"].to_dict())
print("posterior mean mu %.3f sigma %.3f" % (m.output[m.output.phase == "sampling"]["mu"].mean(), m.output[m.output.phase == "sampling"]["sigma"].mean()))
