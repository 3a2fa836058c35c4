/*
 * TEST INFRASTRUCTURE ONLY. Minimal stand-in for libcheck's <check.h> (not installed in this image) so that the
 * reference's own test/test_*.c compile unmodified — against the reference build (oracle check) or against
 * libsdrmodem_b200.so (drop-in proof, oracle/Makefile target dropin-tests). Covers exactly what those files use:
 * START_TEST/END_TEST, ck_assert*, Suite/TCase/SRunner with checked fixtures, CK_NOFORK. A failed assertion prints
 * file:line, marks the test failed and leaves the test function (longjmp), like check does in no-fork mode.
 */
#ifndef SDRM_CHECK_SHIM_H
#define SDRM_CHECK_SHIM_H

#include <setjmp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef void (*ck_test_fn)(int);
typedef void (*ck_fixture_fn)(void);

typedef struct TCase {
    const char *name;
    ck_test_fn tests[128];
    const char *test_names[128];
    int n_tests;
    ck_fixture_fn setup;
    ck_fixture_fn teardown;
} TCase;

typedef struct Suite {
    const char *name;
    TCase *cases[16];
    int n_cases;
} Suite;

typedef struct SRunner {
    Suite *suite;
    int failed;
    int run;
} SRunner;

enum fork_status { CK_FORK_GETENV, CK_FORK, CK_NOFORK };
enum print_output { CK_SILENT, CK_MINIMAL, CK_NORMAL, CK_VERBOSE, CK_ENV, CK_LAST };

static jmp_buf ck_shim_jump;
static int ck_shim_current_failed;
static long ck_shim_checks;

#define START_TEST(testname) static void testname(int _i) { (void) _i;
#define END_TEST }

#define ck_shim_fail(...)                                       \
    do {                                                        \
        fprintf(stderr, "%s:%d: ", __FILE__, __LINE__);         \
        fprintf(stderr, __VA_ARGS__);                           \
        fputc('\n', stderr);                                    \
        ck_shim_current_failed = 1;                             \
        longjmp(ck_shim_jump, 1);                               \
    } while (0)

#define ck_assert(expr)                                         \
    do {                                                        \
        ck_shim_checks++;                                       \
        if (!(expr)) ck_shim_fail("assertion '%s' failed", #expr); \
    } while (0)

#define ck_assert_int_eq(a, b)                                  \
    do {                                                        \
        intmax_t ck_a_ = (intmax_t) (a);                        \
        intmax_t ck_b_ = (intmax_t) (b);                        \
        ck_shim_checks++;                                       \
        if (ck_a_ != ck_b_) ck_shim_fail("%s == %s failed: %jd != %jd", #a, #b, ck_a_, ck_b_); \
    } while (0)

#define ck_assert_uint_eq(a, b)                                 \
    do {                                                        \
        uintmax_t ck_a_ = (uintmax_t) (a);                      \
        uintmax_t ck_b_ = (uintmax_t) (b);                      \
        ck_shim_checks++;                                       \
        if (ck_a_ != ck_b_) ck_shim_fail("%s == %s failed: %ju != %ju", #a, #b, ck_a_, ck_b_); \
    } while (0)

#define ck_assert_str_eq(a, b)                                  \
    do {                                                        \
        ck_shim_checks++;                                       \
        if (strcmp((a), (b)) != 0) ck_shim_fail("%s == %s failed", #a, #b); \
    } while (0)

static inline Suite *suite_create(const char *name) {
    Suite *s = calloc(1, sizeof(Suite));
    s->name = name;
    return s;
}

static inline TCase *tcase_create(const char *name) {
    TCase *t = calloc(1, sizeof(TCase));
    t->name = name;
    return t;
}

static inline void ck_shim_add_test(TCase *tc, ck_test_fn fn, const char *name) {
    tc->tests[tc->n_tests] = fn;
    tc->test_names[tc->n_tests] = name;
    tc->n_tests++;
}
#define tcase_add_test(tc, fn) ck_shim_add_test((tc), (fn), #fn)

static inline void tcase_add_checked_fixture(TCase *tc, ck_fixture_fn setup, ck_fixture_fn teardown) {
    tc->setup = setup;
    tc->teardown = teardown;
}

static inline void suite_add_tcase(Suite *s, TCase *tc) { s->cases[s->n_cases++] = tc; }

static inline SRunner *srunner_create(Suite *s) {
    SRunner *r = calloc(1, sizeof(SRunner));
    r->suite = s;
    return r;
}

static inline void srunner_set_fork_status(SRunner *r, enum fork_status status) {
    (void) r;
    (void) status;
}

static inline void srunner_run_all(SRunner *r, enum print_output mode) {
    (void) mode;
    for (int c = 0; c < r->suite->n_cases; c++) {
        TCase *tc = r->suite->cases[c];
        for (int t = 0; t < tc->n_tests; t++) {
            ck_shim_current_failed = 0;
            if (tc->setup != NULL) tc->setup();
            if (setjmp(ck_shim_jump) == 0) {
                tc->tests[t](0);
            }
            /* checked fixtures: teardown assertions count too */
            if (tc->teardown != NULL) {
                if (setjmp(ck_shim_jump) == 0) {
                    tc->teardown();
                }
            }
            r->run++;
            if (ck_shim_current_failed) {
                r->failed++;
                fprintf(stderr, "%s:%s:%s: FAILED\n", r->suite->name, tc->name, tc->test_names[t]);
            }
        }
    }
    printf("%s: %d tests, %d failed, %ld checks\n", r->suite->name, r->run, r->failed, ck_shim_checks);
}

static inline int srunner_ntests_failed(SRunner *r) { return r->failed; }

static inline void srunner_free(SRunner *r) {
    for (int c = 0; c < r->suite->n_cases; c++) {
        free(r->suite->cases[c]);
    }
    free(r->suite);
    free(r);
}

#endif
