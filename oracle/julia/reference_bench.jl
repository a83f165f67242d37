# reference_bench.jl -- NOT RUN HERE (no Julia in the image).  Times the real reference on BASELINE config 2's
# shape, one signal per call as the reference API works, for a true CPU baseline next to bench.py's numbers.
#   julia --project=/path/to/CompressedSensing.jl -t auto oracle/julia/reference_bench.jl
using CompressedSensing: omp, sparse_data
using BenchmarkTools, LinearAlgebra
M, N, k = 1024, 8192, 32
A, x0, b = sparse_data(n = M, m = N, k = k)
t = @belapsed omp($A, $b, $k)
println("omp 1024x8192 k=32 Float64: ", t * 1e3, " ms per solve = ", 1 / t, " solves/s  (BLAS threads: ", BLAS.get_num_threads(), ")")
