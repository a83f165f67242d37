"""Writes the golden fixtures as raw column-major little-endian .bin + .json (for oracle/julia/reference_oracle.jl)."""
import glob, json, os
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
out = os.path.join(HERE, "bin")
os.makedirs(out, exist_ok=True)
for f in sorted(glob.glob(os.path.join(HERE, "*.npz"))):
    z = np.load(f)
    name = os.path.basename(f)[:-4]
    np.asfortranarray(z["A"]).T.tofile(os.path.join(out, name + ".A.bin"))     # .T of F-order = C-order view of the same bytes
    np.asfortranarray(z["B"]).T.tofile(os.path.join(out, name + ".B.bin"))
    open(os.path.join(out, name + ".json"), "w").write(str(z["meta"]))
print("wrote", out)
