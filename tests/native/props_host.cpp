// Host build of multiprime_b200/csrc/mpb_walk.cu (pure host code) with the device call mpb_tm replaced by a stub
// that records the expansion buffer it is handed: lets the CPU suite check the expansion order of mpb_primer_props.
#include "../../multiprime_b200/csrc/mpb_walk.cu"

static std::vector<uint8_t> g_seqs;
static int64_t g_n = 0;

int mpb_fail(int code, const char*, ...) { return code; }

extern "C" int mpb_tm(mpb_ctx*, const uint8_t* seqs, int k, int64_t n, const double*, double* tm, double*, double*) {
    g_seqs.assign(seqs, seqs + n * k);
    g_n = n;
    for (int64_t i = 0; i < n; ++i) tm[i] = 50.0 + (double)(i % 7);
    return 0;
}

// runs mpb_primer_props and returns the number of expansion rows handed to mpb_tm (copied to out, up to cap rows)
// (gc, flags: the GC content and filter flags it computed; the Tm mean is meaningless here)
extern "C" int64_t props_expansions(const uint8_t* sets, int k, int32_t n, uint8_t* out, int64_t cap, int32_t* deg,
                                    double* gc, int32_t* flags) {
    std::vector<double> tm(n);
    std::vector<int32_t> ndeg(n);
    const double consts[3] = {0, 0, 0};
    int rc = mpb_primer_props(reinterpret_cast<mpb_ctx*>(&g_n), sets, k, n, 0.4, 0.6, 4, consts, tm.data(), gc, flags, deg,
                              ndeg.data());
    if (rc) return rc;
    const int64_t take = g_n < cap ? g_n : cap;
    memcpy(out, g_seqs.data(), (size_t)take * k);
    return g_n;
}
