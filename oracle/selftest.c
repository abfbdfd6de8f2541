/* Sanitizer self-test of the C oracle (test infrastructure): compiled TOGETHER with ragfin_oracle.c under
 * -fsanitize=address,undefined by `make -C oracle selftest` and run by tests/test_oracle_cpu.py.  It exercises the
 * edge cases the parity suites rely on - ragged dims, zero rows, duplicates, k > n, n = 0, every storage dtype, the
 * threaded paths - so that an out-of-bounds access or undefined behaviour in the checker itself cannot hide. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

void oracle_set_threads(int t);
void oracle_synth_rows(uint64_t seed, int64_t row0, int64_t n, int32_t dim, int32_t dup_every, int32_t zero_every, float* out);
void oracle_normalize_rows(const float* x, int64_t n, int32_t dim, int32_t dtype, float* out);
void oracle_exact_scores(const float* stored, int64_t n, int32_t dim, const float* qhat, float* scores);
int oracle_topk(const float* stored, int64_t n, int32_t dim, const float* queries, int32_t nq, int32_t k, int64_t id_base,
                int64_t* ids, float* scores);

static int fails = 0;
#define CHECK(c) do { if (!(c)) { printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); ++fails; } } while (0)

static void one_case(int64_t n, int32_t dim, int32_t nq, int32_t k, int32_t dtype, int threads) {
    oracle_set_threads(threads);
    float* x = (float*)malloc((size_t)(n > 0 ? n : 1) * dim * sizeof(float));
    float* st = (float*)malloc((size_t)(n > 0 ? n : 1) * dim * sizeof(float));
    float* q = (float*)malloc((size_t)nq * dim * sizeof(float));
    int64_t* ids = (int64_t*)malloc((size_t)nq * k * sizeof(int64_t));
    float* sc = (float*)malloc((size_t)nq * k * sizeof(float));
    oracle_synth_rows(11, 5, n, dim, 7, 13, x);
    oracle_synth_rows(12, 0, nq, dim, 0, 0, q);
    oracle_normalize_rows(x, n, dim, dtype, st);
    CHECK(oracle_topk(st, n, dim, q, nq, k, 1000, ids, sc) == 0);
    const int64_t m = n < k ? n : k;
    for (int32_t i = 0; i < nq; ++i) {
        for (int64_t j = 0; j < k; ++j) {
            const int64_t id = ids[(int64_t)i * k + j];
            const float s = sc[(int64_t)i * k + j];
            if (j < m) {
                CHECK(id >= 1000 && id < 1000 + n);
                CHECK(s >= -1.0001f && s <= 1.0001f);
                if (j > 0) {   /* descending score, ties to the lower id */
                    const float sp = sc[(int64_t)i * k + j - 1];
                    CHECK(sp > s || (sp == s && ids[(int64_t)i * k + j - 1] < id));
                }
            } else {
                CHECK(id == -1 && isinf(s) && s < 0);
            }
        }
    }
    free(x); free(st); free(q); free(ids); free(sc);
}

int main(void) {
    const int32_t dims[] = {1, 8, 33, 100, 384, 768};
    for (int d = 0; d < 6; ++d)
        for (int32_t dtype = 0; dtype < 3; ++dtype) {
            one_case(0, dims[d], 2, 3, dtype, 1);        /* empty collection: all padding */
            one_case(5, dims[d], 3, 16, dtype, 1);       /* k > n */
            one_case(257, dims[d], 4, 10, dtype, 3);     /* duplicates (every 7th) and zero rows (every 13th), threaded */
        }
    one_case(20000, 64, 2, 100, 1, 8);
    /* a zero query scores 0 against everything and returns the k lowest ids */
    {
        float st[4 * 8], q[8] = {0}, sc[3];
        int64_t ids[3];
        float x[4 * 8];
        oracle_synth_rows(3, 0, 4, 8, 0, 0, x);
        oracle_normalize_rows(x, 4, 8, 0, st);
        CHECK(oracle_topk(st, 4, 8, q, 1, 3, 0, ids, sc) == 0);
        CHECK(ids[0] == 0 && ids[1] == 1 && ids[2] == 2 && sc[0] == 0.0f && sc[2] == 0.0f);
    }
    printf(fails ? "ORACLE SELFTEST FAILED (%d)\n" : "ORACLE SELFTEST OK\n", fails);
    return fails ? 1 : 0;
}
