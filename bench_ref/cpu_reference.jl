# cpu_reference.jl -- the UNMODIFIED reference (DSP.jl + FFTW.jl) on the headline step of bench.py, timed on the host cores.
#
#   julia [-t N] --project=<environment in which `using DSP` works> bench_ref/cpu_reference.jl <log2n> <steps> <warmup> [fftw_threads]
#
# One step = y = conv(x, v)  (4097-tap complex FIR; conv picks overlap-save with optimalfftfiltlength, src/dspbase.jl:709-792)
#            P = welch_pgram(y[1:n], 4096, 2048; onesided=false, nfft=4096, window=hanning)  (src/periodograms.jl:647-759)
# on n = 2^log2n ComplexF32 samples -- the same taps, window and stage split as bench.py's GPU arm.  Prints ONE JSON line
# that bench.py --impl reference re-emits (cpu_baseline.kind = "reference").  This image has no Julia; the script is
# for boxes that do.  Nothing here is part of the product.
using DSP, FFTW, Random, Printf

function main()
    log2n   = length(ARGS) >= 1 ? parse(Int, ARGS[1]) : 26
    steps   = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 3
    warmup  = length(ARGS) >= 3 ? parse(Int, ARGS[3]) : 1
    threads = length(ARGS) >= 4 ? parse(Int, ARGS[4]) : Sys.CPU_THREADS
    FFTW.set_num_threads(threads)
    n  = 1 << log2n
    nv = 4097
    k  = collect(0:nv-1) .- (nv ÷ 2)
    hamming_w = 0.54 .- 0.46 .* cos.(2pi .* collect(0:nv-1) ./ (nv - 1))             # numpy.hamming
    v  = ComplexF32.(0.2 .* sinc.(0.2 .* k) .* hamming_w .* cis.(pi * 0.3 .* k))
    rng = MersenneTwister(1002)
    x  = ComplexF32.((randn(rng, n) .+ im .* randn(rng, n)) ./ sqrt(2))
    step() = begin
        y = conv(x, v)
        welch_pgram(view(y, 1:n), 4096, 2048; onesided=false, nfft=4096, window=hanning)
    end
    for _ in 1:warmup
        step()
    end
    ts = Float64[]
    for _ in 1:steps
        t0 = time_ns()
        step()
        push!(ts, (time_ns() - t0) * 1e-9)
    end
    dt = sum(ts) / length(ts)
    @printf("{\"impl\": \"reference\", \"kind\": \"reference\", \"value\": %.6g, \"unit\": \"Gsamples/s\", \"ms_per_step\": %.3f, \"cores\": %d, \"julia_threads\": %d, \"sample\": \"2^%d ComplexF32 samples per step; DSP.jl %s conv + welch_pgram, FFTW %d threads\"}\n",
            n / dt / 1e9, dt * 1e3, threads, Threads.nthreads(), log2n, string(pkgversion(DSP)), threads)
end

main()
