# reference_oracle.jl -- TEST INFRASTRUCTURE, NOT RUN IN THIS REPOSITORY'S ENVIRONMENT (no Julia in the image).
#
# For anyone with Julia >= 1.7 and CompressedSensing.jl's pinned manifest: runs the REAL reference on the same
# bytes as the committed fixtures and prints what tests/golden/*.npz store, so that the oracle
# (oracle/pursuit_oracle.py) and the CUDA path can be pinned against the true reference.
#
#   python tests/golden/export_bin.py            # writes tests/golden/bin/<name>.{A,B}.bin + <name>.json
#   julia --project=/path/to/CompressedSensing.jl oracle/julia/reference_oracle.jl tests/golden/bin
#
# Output: one line per signal:  <fixture> <signal> nzind=[...] nzval=[...] resnorm=...   (1-based indices)
using CompressedSensing: omp, gomp, mp
using LinearAlgebra, SparseArrays, JSON

dir = length(ARGS) >= 1 ? ARGS[1] : "tests/golden/bin"
for meta_file in sort(filter(f -> endswith(f, ".json"), readdir(dir; join = true)))
    meta = JSON.parsefile(meta_file)
    T = meta["dtype"] == "float32" ? Float32 : Float64
    M, N, B, k = meta["M"], meta["N"], meta["B"], meta["k"]
    base = replace(meta_file, ".json" => "")
    A = Matrix{T}(undef, M, N); read!(base * ".A.bin", A)        # column-major, little-endian
    Bm = Matrix{T}(undef, M, B); read!(base * ".B.bin", Bm)
    for s in 1:B
        b = Bm[:, s]
        x = meta["algo"] == "omp" ? omp(A, b, k) :
            meta["algo"] == "gomp" ? gomp(A, b, meta["l"], k) : mp(A, b, k)
        println(basename(base), " ", s - 1, " nzind=", x.nzind, " nzval=", x.nzval, " resnorm=", norm(b - A * x))
    end
end
