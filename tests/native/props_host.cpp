// Host build of multiprime_b200/csrc/mpb_walk.cu (pure host code) with the device call mpb_tm_sets replaced by a stub:
// lets the CPU suite check what mpb_primer_props computes on the host (degeneracy, GC content, di-nucleotide / hairpin
// flags, the rounding of the Tm mean).
#include "../../multiprime_b200/csrc/mpb_walk.cu"

int mpb_fail(int code, const char*, ...) { return code; }

// every expansion "has" Tm 50.00 + 0.01 * (its primer index % 7): sums are then known in closed form
extern "C" int mpb_tm_sets(mpb_ctx*, const uint8_t* sets, int k, int32_t n, const double*, int64_t* sums, int32_t* ties) {
    for (int i = 0; i < n; ++i) {
        int nd = 0;
        sums[i] = (int64_t)degeneracy_of(sets + (int64_t)i * 32, k, &nd) * (5000 + i % 7);
        ties[i] = 0;
    }
    return 0;
}

extern "C" int props_host(const uint8_t* sets, int k, int32_t n, int32_t* deg, int32_t* ndeg, double* tm, double* gc,
                          int32_t* flags) {
    const double consts[3] = {0, 0, 0};
    int dummy = 0;
    return mpb_primer_props(reinterpret_cast<mpb_ctx*>(&dummy), sets, k, n, 0.4, 0.6, 4, consts, tm, gc, flags, deg, ndeg);
}
