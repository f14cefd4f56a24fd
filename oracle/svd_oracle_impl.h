/* TEST INFRASTRUCTURE ONLY -- type-generic body of the CPU oracle (included once per element type).
 *
 * Plain-C restatement of the reference's two-stage bidiagonal reduction + QR diagonalisation,
 * following the reference's arithmetic ORDER (k-ascending sums that start from 0, separate
 * multiply and add -- compile with -ffp-contract=off --, reflector scalars evaluated in double)
 * so that it reproduces the reference bit-for-bit.  Pinned against data/band_* and
 * data/bidiagonal_* (tests/test_oracle.py) and against oracle/_ref (the compiled reference).
 *
 * Required macros: T (element type), FN(name) (symbol suffixing), SQRT_T, FABS_T.
 * All matrices are dense row-major with explicit leading dimensions.
 */

/* C = A(m x kk) * B(kk x n); matrix.h:234-248: result starts at 0, k ascending, acc += a*b in T. */
static void FN(mm)(T* C, size_t ldc, const T* A, size_t lda, const T* B, size_t ldb,
                   size_t m, size_t kk, size_t n) {
    for (size_t i = 0; i < m; ++i)
        for (size_t j = 0; j < n; ++j) {
            T acc = 0;
            for (size_t k = 0; k < kk; ++k) acc += A[i * lda + k] * B[k * ldb + j];
            C[i * ldc + j] = acc;
        }
}

/* svd_serial.h:189-216.  x has length len (stride incx).  Writes w (len) and tau.
 * s, u1, 1/u1 and tau are evaluated in double even for T=float (copysign(int,T) promotes),
 * the norm is accumulated in T in index order (matrix.h:59-62). */
static void FN(householder)(const T* x, size_t incx, size_t len, T* w, T* tau) {
    double s = -copysign(1.0, (double)x[0]);
    T acc = 0;
    for (size_t i = 0; i < len; ++i) acc = acc + x[i * incx] * x[i * incx];
    T norm_x = SQRT_T(acc);
    double u1 = (double)x[0] - s * (double)norm_x;
    T alpha = (T)(1. / u1);
    for (size_t i = 0; i < len; ++i) w[i] = x[i * incx] * alpha;
    w[0] = (T)1.;
    *tau = (T)(-s * u1 / (double)norm_x);
}

/* Explicit H = I - tau w w^T exactly as svd_serial.h:204-211: H_ij = (0 + w_i*w_j) * (-tau); H_dd = 1 + H_dd. */
static void FN(hh_transform)(const T* w, size_t len, T tau, T* H) {
    T mt = -tau;
    for (size_t i = 0; i < len; ++i)
        for (size_t j = 0; j < len; ++j) {
            T p = 0;
            p += w[i] * w[j];
            H[i * len + j] = p * mt;
        }
    for (size_t d = 0; d < len; ++d) H[d * len + d] = 1 + H[d * len + d];
}

/* svd_parallel.h:97-113.  V is (vrows x t, ldv); S is t x t (lds). Column j of V already holds v_j. */
static void FN(hholder_compact)(size_t j, T tau, T* S, size_t lds, const T* V, size_t ldv, size_t vrows,
                                T* z, T* z2) {
    if (j == 0) { S[0] = -tau; return; }
    for (size_t r = 0; r < j; ++r) {            /* z = V_k^T v : sum over ALL rows of V, ascending */
        T acc = 0;
        for (size_t i = 0; i < vrows; ++i) acc += V[i * ldv + r] * V[i * ldv + j];
        z[r] = acc;
    }
    for (size_t r = 0; r < j; ++r) {            /* z2 = S_k z */
        T acc = 0;
        for (size_t c = 0; c < j; ++c) acc += S[r * lds + c] * z[c];
        z2[r] = acc;
    }
    T mt = -tau;
    for (size_t r = 0; r < j; ++r) S[r * lds + j] = z2[r] * mt;
    S[j * lds + j] = -tau;
}

/* svd_parallel.h:133-169.  A: m x n (lda) overwritten by R; S: n x n persistent; V: m x n persistent. */
static void FN(qr)(T* A, size_t lda, size_t m, size_t n, T* S, size_t lds, T* V, size_t ldv, T* work) {
    T* Y = work;             /* n x n, fresh zeros */
    T* R = Y + n * n;        /* m x n */
    T* w = R + m * n;        /* m */
    T* z = w + m;            /* n */
    T* z2 = z + n;           /* n */
    T* P = z2 + n;           /* m x n : V*Y */
    memset(Y, 0, sizeof(T) * n * n);
    size_t kmax = n < m ? n : m;
    for (size_t j = 0; j < kmax; ++j) {
        FN(mm)(P, n, V, ldv, Y, n, m, n, n);
        for (size_t i = 0; i < m; ++i)
            for (size_t c = 0; c < n; ++c) R[i * n + c] = A[i * lda + c] - P[i * n + c];
        T tau;
        FN(householder)(R + j * n + j, n, m - j, w, &tau);
        for (size_t c = j; c < n; ++c) {        /* y = tau * R[j:,j:]^T w */
            T acc = 0;
            for (size_t r = 0; r < m - j; ++r) acc += R[(j + r) * n + c] * w[r];
            Y[j * n + c] = acc * tau;
        }
        for (size_t r = 0; r < m - j; ++r) V[(j + r) * ldv + j] = w[r];
        FN(hholder_compact)(j, tau, S, lds, V, ldv, m, z, z2);
    }
    FN(mm)(P, n, V, ldv, Y, n, m, n, n);
    for (size_t i = 0; i < m; ++i)
        for (size_t c = 0; c < n; ++c) A[i * lda + c] = A[i * lda + c] - P[i * n + c];
}

/* svd_parallel.h:189-226.  A: m x n (m <= n) overwritten by L; S: m x m persistent; U: m x n persistent. */
static void FN(lq)(T* A, size_t lda, size_t m, size_t n, T* S, size_t lds, T* U, size_t ldu, T* work) {
    T* X = work;             /* m x m fresh zeros */
    T* L = X + m * m;        /* m x n */
    T* w = L + m * n;        /* n */
    T* z = w + n;            /* m */
    T* z2 = z + m;           /* m */
    T* P = z2 + m;           /* m x n : X*U */
    T* UT = P + m * n;       /* n x m */
    memset(X, 0, sizeof(T) * m * m);
    size_t kmax = n < m ? n : m;
    for (size_t i = 0; i < kmax; ++i) {
        FN(mm)(P, n, X, m, U, ldu, m, m, n);
        for (size_t r = 0; r < m; ++r)
            for (size_t c = 0; c < n; ++c) L[r * n + c] = A[r * lda + c] - P[r * n + c];
        T tau;
        FN(householder)(L + i * n + i, 1, n - i, w, &tau);
        for (size_t r = i; r < m; ++r) {        /* x = tau * L[i:,i:] w */
            T acc = 0;
            for (size_t c = 0; c < n - i; ++c) acc += L[r * n + i + c] * w[c];
            X[r * m + i] = acc * tau;
        }
        for (size_t c = 0; c < n - i; ++c) U[i * ldu + i + c] = w[c];
        for (size_t r = 0; r < m; ++r)
            for (size_t c = 0; c < n; ++c) UT[c * m + r] = U[r * ldu + c];
        FN(hholder_compact)(i, tau, S, lds, UT, m, n, z, z2);
    }
    FN(mm)(P, n, X, m, U, ldu, m, m, n);
    for (size_t r = 0; r < m; ++r)
        for (size_t c = 0; c < n; ++c) A[r * lda + c] = A[r * lda + c] - P[r * n + c];
}

/* Qt = (V (S V^T))  stored so that apply uses Q2[k][i]; svd_parallel.h:243-254 (first four lines). */
static void FN(form_q)(T* Q2, const T* S, size_t lds, const T* V, size_t ldv, size_t rows, size_t t, T* Q1) {
    for (size_t r = 0; r < t; ++r)              /* Q1 = S * V^T : t x rows */
        for (size_t c = 0; c < rows; ++c) {
            T acc = 0;
            for (size_t k = 0; k < t; ++k) acc += S[r * lds + k] * V[c * ldv + k];
            Q1[r * rows + c] = acc;
        }
    FN(mm)(Q2, rows, V, ldv, Q1, rows, rows, t, rows);   /* Q2 = V * Q1 : rows x rows */
}
/* A(rows x cols) += Q2^T A ; svd_parallel.h:250-253 */
static void FN(qr_apply_q)(T* A, size_t lda, size_t rows, size_t cols, const T* Q2, T* P) {
    for (size_t i = 0; i < rows; ++i)
        for (size_t c = 0; c < cols; ++c) {
            T acc = 0;
            for (size_t k = 0; k < rows; ++k) acc += Q2[k * rows + i] * A[k * lda + c];
            P[i * cols + c] = acc;
        }
    for (size_t i = 0; i < rows; ++i)
        for (size_t c = 0; c < cols; ++c) A[i * lda + c] = A[i * lda + c] + P[i * cols + c];
}
/* P = U^T (S U) : cols x cols ; svd_parallel.h:271-278 */
static void FN(form_p)(T* Pm, const T* S, size_t lds, const T* U, size_t ldu, size_t cols, size_t t, T* P1) {
    FN(mm)(P1, cols, S, lds, U, ldu, t, t, cols);        /* P1 = S * U : t x cols */
    for (size_t i = 0; i < cols; ++i)
        for (size_t c = 0; c < cols; ++c) {
            T acc = 0;
            for (size_t k = 0; k < t; ++k) acc += U[k * ldu + i] * P1[k * cols + c];
            Pm[i * cols + c] = acc;
        }
}
/* A(rows x cols) += A P ; svd_parallel.h:280 */
static void FN(lq_apply_p)(T* A, size_t lda, size_t rows, size_t cols, const T* Pm, T* R) {
    FN(mm)(R, cols, A, lda, Pm, cols, rows, cols, cols);
    for (size_t i = 0; i < rows; ++i)
        for (size_t c = 0; c < cols; ++c) A[i * lda + c] = A[i * lda + c] + R[i * cols + c];
}

/* Generic (un-hoisted) application kernels, exported for unit parity tests against the reference. */
int FN(svdo_qr_apply)(T* A, size_t rows, size_t cols, const T* S, const T* V, size_t t) {
    T* Q1 = (T*)malloc(sizeof(T) * (t * rows + rows * rows + rows * cols));
    T* Q2 = Q1 + t * rows; T* P = Q2 + rows * rows;
    FN(form_q)(Q2, S, t, V, t, rows, t, Q1);
    FN(qr_apply_q)(A, cols, rows, cols, Q2, P);
    free(Q1);
    return 0;
}
int FN(svdo_lq_apply)(T* A, size_t rows, size_t cols, const T* S, const T* U, size_t t) {
    T* P1 = (T*)malloc(sizeof(T) * (t * cols + cols * cols + rows * cols));
    T* Pm = P1 + t * cols; T* R = Pm + cols * cols;
    FN(form_p)(Pm, S, t, U, cols, cols, t, P1);
    FN(lq_apply_p)(A, cols, rows, cols, Pm, R);
    free(P1);
    return 0;
}
int FN(svdo_panel_qr)(T* A, size_t m, size_t n, T* S, T* V) {
    T* work = (T*)calloc(n * n + 2 * m * n + m + 2 * n + 16, sizeof(T));
    FN(qr)(A, n, m, n, S, n, V, n, work);
    free(work);
    return 0;
}
int FN(svdo_panel_lq)(T* A, size_t m, size_t n, T* S, T* U) {
    T* work = (T*)calloc(m * m + 3 * m * n + n + 2 * m + 16, sizeof(T));
    FN(lq)(A, n, m, n, S, m, U, n, work);
    free(work);
    return 0;
}
/* csc586::serial::brd<T>, svd_serial.h:233-266: one-stage Golub-Kahan bidiagonalisation.  For every column j: explicit H
 * of A[j:m, j] applied from the left to A[j:m, j:n] (H.mm(minor)); then, for j < n-1, explicit H of the row A[j, j+1:n]
 * applied from the right to A[j:m, j+1:n] (minor.mm(H)).  Length-1 reflectors flip a sign (tau = 2). */
int FN(svdo_brd_serial)(T* A, size_t n, T* d, T* e) {
    T* w = (T*)malloc(sizeof(T) * n);
    T* H = (T*)malloc(sizeof(T) * n * n);
    T* minor = (T*)malloc(sizeof(T) * n * n);
    T* prod = (T*)malloc(sizeof(T) * n * n);
    if (!w || !H || !minor || !prod) { free(w); free(H); free(minor); free(prod); return 1; }
    for (size_t j = 0; j < n; ++j) {
        T tau;
        size_t len = n - j, cols = n - j;
        FN(householder)(A + j * n + j, n, len, w, &tau);
        FN(hh_transform)(w, len, tau, H);
        for (size_t r = 0; r < len; ++r) memcpy(minor + r * cols, A + (j + r) * n + j, sizeof(T) * cols);
        FN(mm)(prod, cols, H, len, minor, cols, len, len, cols);
        for (size_t r = 0; r < len; ++r) memcpy(A + (j + r) * n + j, prod + r * cols, sizeof(T) * cols);
        if (j + 1 < n) {
            size_t rl = n - j - 1;                       /* reflector length = columns j+1 .. n-1 */
            FN(householder)(A + j * n + j + 1, 1, rl, w, &tau);
            FN(hh_transform)(w, rl, tau, H);
            for (size_t r = 0; r < len; ++r) memcpy(minor + r * rl, A + (j + r) * n + j + 1, sizeof(T) * rl);
            FN(mm)(prod, rl, minor, rl, H, rl, len, rl, rl);
            for (size_t r = 0; r < len; ++r) memcpy(A + (j + r) * n + j + 1, prod + r * rl, sizeof(T) * rl);
        }
    }
    for (size_t i = 0; i < n; ++i) { if (d) d[i] = A[i * n + i]; if (e && i + 1 < n) e[i] = A[i * n + i + 1]; }
    free(w); free(H); free(minor); free(prod);
    return 0;
}

int FN(svdo_householder)(const T* x, size_t len, T* w, T* H, T* tau) {
    FN(householder)(x, 1, len, w, tau);
    if (H) FN(hh_transform)(w, len, *tau, H);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Stage 1, tile flat-tree order: csc586::parallel::brd_p1<T>, svd_parallel.h:411-533.
 * Square n x n, t | n.  In place.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    size_t t;
    T *S_kk, *V_kk, *S_ik, *V_ik, *S_ki, *V_ki, *S_kk1, *V_kk1, *stack, *work;
} FN(tilectx);

/* factor_2tile (svd_parallel.h:311-339) for the row-wise (QR) case: [tile(k,k); tile(i,k)] */
static void FN(ts_qr)(T* A, size_t n, size_t t, size_t k, size_t i, FN(tilectx)* c) {
    T* st = c->stack;                                   /* 2t x t */
    for (size_t r = 0; r < t; ++r) {
        memcpy(st + r * t, A + (k * t + r) * n + k * t, sizeof(T) * t);
        memcpy(st + (t + r) * t, A + (i * t + r) * n + k * t, sizeof(T) * t);
    }
    FN(qr)(st, t, 2 * t, t, c->S_ik, t, c->V_ik, t, c->work);
    for (size_t r = 0; r < t; ++r) {
        memcpy(A + (k * t + r) * n + k * t, st + r * t, sizeof(T) * t);
        memcpy(A + (i * t + r) * n + k * t, st + (t + r) * t, sizeof(T) * t);
    }
}
/* factor_2tile column-wise (LQ) case: [tile(k,k+1) | tile(k,i)] */
static void FN(ts_lq)(T* A, size_t n, size_t t, size_t k, size_t i, FN(tilectx)* c) {
    T* st = c->stack;                                   /* t x 2t */
    for (size_t r = 0; r < t; ++r) {
        memcpy(st + r * 2 * t, A + (k * t + r) * n + (k + 1) * t, sizeof(T) * t);
        memcpy(st + r * 2 * t + t, A + (k * t + r) * n + i * t, sizeof(T) * t);
    }
    FN(lq)(st, 2 * t, t, 2 * t, c->S_ki, t, c->V_ki, 2 * t, c->work);
    for (size_t r = 0; r < t; ++r) {
        memcpy(A + (k * t + r) * n + (k + 1) * t, st + r * 2 * t, sizeof(T) * t);
        memcpy(A + (k * t + r) * n + i * t, st + r * 2 * t + t, sizeof(T) * t);
    }
}

int FN(svdo_brd_p1)(T* A, size_t n, size_t t) {
    if (t == 0 || n == 0 || n % t != 0) return -1;
    size_t nbt = n / t;
    size_t wsz = 16 * t * t + 8 * t + 64;
    FN(tilectx) c;
    c.t = t;
    T* pool = (T*)calloc(t * t * 4 + 2 * t * t * 2 + t * t * 2 + 2 * t * t + wsz, sizeof(T));
    if (!pool) return -2;
    c.S_kk = pool; c.V_kk = c.S_kk + t * t; c.S_ik = c.V_kk + t * t; c.V_ik = c.S_ik + t * t;
    c.S_ki = c.V_ik + 2 * t * t; c.V_ki = c.S_ki + t * t; c.S_kk1 = c.V_ki + 2 * t * t;
    c.V_kk1 = c.S_kk1 + t * t; c.stack = c.V_kk1 + t * t; c.work = c.stack + 2 * t * t;
    T* Q = (T*)malloc(sizeof(T) * (4 * t * t + 2 * t * t));   /* Q2 (2t x 2t) + scratch */
    T* Qs = Q + 4 * t * t;

    for (size_t k = 0; k < nbt; ++k) {
        T* Akk = A + (k * t) * n + k * t;
        /* QR step 1 (443-445) */
        if (k == 0 || k == nbt - 1) FN(qr)(Akk, n, t, t, c.S_kk, t, c.V_kk, t, c.work);
        /* QR step 2 (452-461): apply (S_kk,V_kk) along tile row k */
        if (k + 1 < nbt) {
            FN(form_q)(Q, c.S_kk, t, c.V_kk, t, t, t, Qs);
            #pragma omp parallel
            {
                T* P = (T*)malloc(sizeof(T) * t * t);
                #pragma omp for schedule(static)
                for (size_t j = k + 1; j < nbt; ++j)
                    FN(qr_apply_q)(A + (k * t) * n + j * t, n, t, t, Q, P);
                free(P);
            }
        }
        /* QR steps 3+4 (466-486) */
        for (size_t i = k + 1; i < nbt; ++i) {
            FN(ts_qr)(A, n, t, k, i, &c);
            FN(form_q)(Q, c.S_ik, t, c.V_ik, t, 2 * t, t, Qs);
            #pragma omp parallel
            {
                T* st = (T*)malloc(sizeof(T) * 4 * t * t);
                T* P = st + 2 * t * t;
                #pragma omp for schedule(static)
                for (size_t j = k + 1; j < nbt; ++j) {
                    for (size_t r = 0; r < t; ++r) {
                        memcpy(st + r * t, A + (k * t + r) * n + j * t, sizeof(T) * t);
                        memcpy(st + (t + r) * t, A + (i * t + r) * n + j * t, sizeof(T) * t);
                    }
                    FN(qr_apply_q)(st, t, 2 * t, t, Q, P);
                    for (size_t r = 0; r < t; ++r) {
                        memcpy(A + (k * t + r) * n + j * t, st + r * t, sizeof(T) * t);
                        memcpy(A + (i * t + r) * n + j * t, st + (t + r) * t, sizeof(T) * t);
                    }
                }
                free(st);
            }
        }
        if (k + 1 >= nbt) break;
        /* LQ of tile (k,k+1) (483) */
        T* Akk1 = A + (k * t) * n + (k + 1) * t;
        FN(lq)(Akk1, n, t, t, c.S_kk1, t, c.V_kk1, t, c.work);
        /* LQ step 2 (499-508): apply along tile column k+1, rows j = k+1.. */
        FN(form_p)(Q, c.S_kk1, t, c.V_kk1, t, t, t, Qs);
        #pragma omp parallel
        {
            T* P = (T*)malloc(sizeof(T) * t * t);
            #pragma omp for schedule(static)
            for (size_t j = k + 1; j < nbt; ++j)
                FN(lq_apply_p)(A + (j * t) * n + (k + 1) * t, n, t, t, Q, P);
            free(P);
        }
        /* LQ steps 3+4 (512-528) */
        for (size_t i = k + 2; i < nbt; ++i) {
            FN(ts_lq)(A, n, t, k, i, &c);
            FN(form_p)(Q, c.S_ki, t, c.V_ki, 2 * t, 2 * t, t, Qs);
            #pragma omp parallel
            {
                T* st = (T*)malloc(sizeof(T) * 4 * t * t);
                T* P = st + 2 * t * t;
                #pragma omp for schedule(static)
                for (size_t j = k + 1; j < nbt; ++j) {
                    for (size_t r = 0; r < t; ++r) {
                        memcpy(st + r * 2 * t, A + (j * t + r) * n + (k + 1) * t, sizeof(T) * t);
                        memcpy(st + r * 2 * t + t, A + (j * t + r) * n + i * t, sizeof(T) * t);
                    }
                    FN(lq_apply_p)(st, 2 * t, t, 2 * t, Q, P);
                    for (size_t r = 0; r < t; ++r) {
                        memcpy(A + (j * t + r) * n + (k + 1) * t, st + r * 2 * t, sizeof(T) * t);
                        memcpy(A + (j * t + r) * n + i * t, st + r * 2 * t + t, sizeof(T) * t);
                    }
                }
                free(st);
            }
        }
        /* QR of the next diagonal tile (525); for k = nbt-2 the LQ i-loop is empty so it is done at 444 */
        if (k + 2 < nbt) {
            T* An = A + ((k + 1) * t) * n + (k + 1) * t;
            FN(qr)(An, n, t, t, c.S_kk, t, c.V_kk, t, c.work);
        }
    }
    free(Q);
    free(pool);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Stage 1, full-height panel order: csc586::gpu::brd_p1 (svd_cpu.h:370-425) == the structure of
 * cuda_brd_p1 (svd_cuda_1.cu:750, svd_cuda_2.cu:1117).  Same primitives, reference order.
 * (The reference's twin is float-only; the double instantiation is the same text with T=double.)
 * ------------------------------------------------------------------------------------------- */
int FN(svdo_brd_p1_panel)(T* A, size_t n, size_t b) {
    if (b == 0 || n == 0 || n % b != 0) return -1;
    T* S = (T*)calloc(b * b, sizeof(T));
    for (size_t k = 0; k < n; k += b) {
        size_t m = n - k;                      /* trailing rows */
        size_t nc = n - k - b;                 /* trailing cols right of the panel */
        /* QR panel: rows k.., cols k..k+b */
        T* V = (T*)calloc(m * b, sizeof(T));
        T* work = (T*)calloc(b * b + 2 * m * b + m + 2 * b + 16, sizeof(T));
        FN(qr)(A + k * n + k, n, m, b, S, b, V, b, work);
        free(work);
        if (nc > 0) {
            /* A_trail += (V S V^T)^T A_trail, forming the m x m Q exactly like svd_cpu.h:345-353 so
             * the float instantiation is bit-pinned against gpu::brd_p1 (test sizes only: O(m^2) memory). */
            T* Q1 = (T*)malloc(sizeof(T) * (b * m + m * m));
            T* Q2 = Q1 + b * m;
            FN(form_q)(Q2, S, b, V, b, m, b, Q1);
            T* P = (T*)malloc(sizeof(T) * m * nc);
            FN(qr_apply_q)(A + k * n + k + b, n, m, nc, Q2, P);
            free(P); free(Q1);
        }
        free(V);
        if (k + b < n - 1) {
            size_t mr = m - b;                 /* rows below the row panel */
            T* U = (T*)calloc(b * nc, sizeof(T));
            T* work2 = (T*)calloc(b * b + 3 * b * nc + nc + 2 * b + 16, sizeof(T));
            FN(lq)(A + k * n + k + b, n, b, nc, S, b, U, nc, work2);
            free(work2);
            if (mr > 0) {
                T* P1 = (T*)malloc(sizeof(T) * (b * nc + nc * nc));
                T* Pm = P1 + b * nc;
                FN(form_p)(Pm, S, b, U, nc, nc, b, P1);
                T* R = (T*)malloc(sizeof(T) * mr * nc);
                FN(lq_apply_p)(A + (k + b) * n + k + b, n, mr, nc, Pm, R);
                free(R); free(P1);
            }
            free(U);
        }
    }
    free(S);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Stage 2: csc586::parallel::brd_p2<T>, svd_parallel.h:640-695 with band_rd_top/right/left
 * (569-624).  Explicit dense H per window, window*H / H*window with k-ascending sums.
 * ------------------------------------------------------------------------------------------- */
static void FN(win_right)(T* A, size_t n, size_t i1, size_t i2, size_t j1, size_t j2, T* w, T* H, T* tmp) {
    size_t nr = i2 - i1, nc = j2 - j1;
    T tau;
    FN(householder)(A + i1 * n + j1, 1, nc, w, &tau);
    FN(hh_transform)(w, nc, tau, H);
    FN(mm)(tmp, nc, A + i1 * n + j1, n, H, nc, nr, nc, nc);
    for (size_t r = 0; r < nr; ++r) memcpy(A + (i1 + r) * n + j1, tmp + r * nc, sizeof(T) * nc);
}
static void FN(win_left)(T* A, size_t n, size_t i1, size_t i2, size_t j1, size_t j2, T* w, T* H, T* tmp) {
    size_t nr = i2 - i1, nc = j2 - j1;
    T tau;
    FN(householder)(A + i1 * n + j1, n, nr, w, &tau);
    FN(hh_transform)(w, nr, tau, H);
    FN(mm)(tmp, nc, H, nr, A + i1 * n + j1, n, nr, nr, nc);
    for (size_t r = 0; r < nr; ++r) memcpy(A + (i1 + r) * n + j1, tmp + r * nc, sizeof(T) * nc);
}
#define SVDO_MIN(a, b) ((a) < (b) ? (a) : (b))
int FN(svdo_brd_p2)(T* A, size_t n, size_t band, T* d, T* e) {
    if (n < 2) return -1;
    size_t m = n, w_ = band + 1;                       /* b_size += 1 (648) */
    T* w = (T*)malloc(sizeof(T) * (2 * w_ + 4 * w_ * w_ + 4 * w_ * w_));
    T* H = w + 2 * w_;
    T* tmp = H + 4 * w_ * w_;
    for (size_t i = 0; i + 1 < n; ++i) {
        size_t end_i = SVDO_MIN(i + w_, m), end_j = SVDO_MIN(i + w_, n);
        size_t li1 = i, li2 = end_i, lj1 = i + 1, lj2 = end_j;   /* t_left (658) */
        /* band_rd_top (569-589) */
        FN(win_right)(A, n, li1, li2, lj1, lj2, w, H, tmp);
        lj2 = SVDO_MIN(i + w_ + w_ - 1, n);
        li1 = li1 + 1; lj1 = i + 1;
        FN(win_left)(A, n, li1, li2, lj1, lj2, w, H, tmp);
        if (w_ < 2) continue;
        size_t nbtx = (n - lj2) / (w_ - 1);            /* integer division; the ceil at 664 is a no-op */
        for (size_t k = 0; k < nbtx + 1; ++k) {
            size_t ei = SVDO_MIN(li2 + w_ - 1, m);
            size_t sj = SVDO_MIN(lj1 + w_ - 1, n);
            size_t ej3 = SVDO_MIN(lj2 + w_ - 1, n);
            size_t ri1 = li1, ri2 = ei, rj1 = sj, rj2 = lj2;     /* t_right (675) */
            li1 = li2; li2 = ei; lj1 = sj; lj2 = ej3;            /* t_left  (676) */
            if (rj2 > rj1) FN(win_right)(A, n, ri1, ri2, rj1, rj2, w, H, tmp);
            if (lj2 > lj1) FN(win_left)(A, n, li1, li2, lj1, lj2, w, H, tmp);
        }
    }
    if (d) for (size_t i = 0; i < n; ++i) d[i] = A[i * n + i];
    if (e) for (size_t i = 0; i + 1 < n; ++i) e[i] = A[i * n + i + 1];
    free(w);
    return 0;
}

/* Window ops of the complete variant: identical to win_right / win_left except that a zero vector (sum of squares == 0,
 * which the extra windows at the matrix edge can meet, and which makes the reference formula divide by zero) is left
 * alone: alpha = tau = 0, i.e. H = I built and applied with the same operations. */
static void FN(win_right_g)(T* A, size_t n, size_t i1, size_t i2, size_t j1, size_t j2, T* w, T* H, T* tmp) {
    size_t nr = i2 - i1, nc = j2 - j1;
    T tau, acc = 0;
    for (size_t i = 0; i < nc; ++i) acc = acc + A[i1 * n + j1 + i] * A[i1 * n + j1 + i];
    if (acc == 0) { for (size_t i = 0; i < nc; ++i) w[i] = A[i1 * n + j1 + i] * (T)0; w[0] = (T)1.; tau = (T)0; }
    else FN(householder)(A + i1 * n + j1, 1, nc, w, &tau);
    FN(hh_transform)(w, nc, tau, H);
    FN(mm)(tmp, nc, A + i1 * n + j1, n, H, nc, nr, nc, nc);
    for (size_t r = 0; r < nr; ++r) memcpy(A + (i1 + r) * n + j1, tmp + r * nc, sizeof(T) * nc);
}
static void FN(win_left_g)(T* A, size_t n, size_t i1, size_t i2, size_t j1, size_t j2, T* w, T* H, T* tmp) {
    size_t nr = i2 - i1, nc = j2 - j1;
    T tau, acc = 0;
    for (size_t i = 0; i < nr; ++i) acc = acc + A[(i1 + i) * n + j1] * A[(i1 + i) * n + j1];
    if (acc == 0) { for (size_t i = 0; i < nr; ++i) w[i] = A[(i1 + i) * n + j1] * (T)0; w[0] = (T)1.; tau = (T)0; }
    else FN(householder)(A + i1 * n + j1, n, nr, w, &tau);
    FN(hh_transform)(w, nr, tau, H);
    FN(mm)(tmp, nc, H, nr, A + i1 * n + j1, n, nr, nr, nc);
    for (size_t r = 0; r < nr; ++r) memcpy(A + (i1 + r) * n + j1, tmp + r * nc, sizeof(T) * nc);
}

/* Stage 2 with a COMPLETE chase (not the reference's schedule): svd_parallel.h:664 computes the number of window pairs
 * of a sweep with an integer division inside ceil(), i.e. floor; when (n - t_left.j2) is not a multiple of w-1 the last
 * LEFT window leaves a bulge that no RIGHT window chases and the result is no longer orthogonally equivalent to the
 * input.  This variant keeps emitting window pairs until both are empty -- same windows otherwise, same
 * arithmetic -- and is the checker for the product's optional complete schedule (svdb200_set_stage2_schedule). */
int FN(svdo_brd_p2_complete)(T* A, size_t n, size_t band, T* d, T* e) {
    if (n < 2) return -1;
    size_t m = n, w_ = band + 1;
    T* w = (T*)malloc(sizeof(T) * (2 * w_ + 4 * w_ * w_ + 4 * w_ * w_));
    T* H = w + 2 * w_;
    T* tmp = H + 4 * w_ * w_;
    for (size_t i = 0; i + 1 < n; ++i) {
        size_t end_i = SVDO_MIN(i + w_, m), end_j = SVDO_MIN(i + w_, n);
        size_t li1 = i, li2 = end_i, lj1 = i + 1, lj2 = end_j;
        FN(win_right_g)(A, n, li1, li2, lj1, lj2, w, H, tmp);
        lj2 = SVDO_MIN(i + w_ + w_ - 1, n);
        li1 = li1 + 1; lj1 = i + 1;
        FN(win_left_g)(A, n, li1, li2, lj1, lj2, w, H, tmp);
        if (w_ < 2) continue;
        for (;;) {
            size_t ei = SVDO_MIN(li2 + w_ - 1, m);
            size_t sj = SVDO_MIN(lj1 + w_ - 1, n);
            size_t ej3 = SVDO_MIN(lj2 + w_ - 1, n);
            size_t ri1 = li1, ri2 = ei, rj1 = sj, rj2 = lj2;
            int did = 0;
            li1 = li2; li2 = ei; lj1 = sj; lj2 = ej3;
            if (rj2 > rj1 && ri2 > ri1) { FN(win_right_g)(A, n, ri1, ri2, rj1, rj2, w, H, tmp); did = 1; }
            if (lj2 > lj1 && li2 > li1) { FN(win_left_g)(A, n, li1, li2, lj1, lj2, w, H, tmp); did = 1; }
            if (!did) break;
        }
    }
    if (d) for (size_t i = 0; i < n; ++i) d[i] = A[i * n + i];
    if (e) for (size_t i = 0; i + 1 < n; ++i) e[i] = A[i * n + i + 1];
    free(w);
    return 0;
}

/* Window schedule only (no arithmetic): fills out[] with 6-tuples {kind(0=right,1=left),i1,i2,j1,j2,sweep}.
 * Returns the number of windows; out may be NULL to count.  Used to pin the closed form in the
 * CUDA kernel (SURVEY 8a'') against the reference recurrence. */
size_t FN(svdo_brd_p2_schedule)(size_t n, size_t band, long long* out, size_t cap) {
    size_t m = n, w_ = band + 1, cnt = 0;
#define EMIT(kind, a, b, c_, d_) do { if (out && cnt < cap) { long long* o = out + 6 * cnt; o[0] = kind; o[1] = (long long)(a); \
        o[2] = (long long)(b); o[3] = (long long)(c_); o[4] = (long long)(d_); o[5] = (long long)i; } ++cnt; } while (0)
    for (size_t i = 0; i + 1 < n; ++i) {
        size_t end_i = SVDO_MIN(i + w_, m), end_j = SVDO_MIN(i + w_, n);
        size_t li1 = i, li2 = end_i, lj1 = i + 1, lj2 = end_j;
        EMIT(0, li1, li2, lj1, lj2);
        lj2 = SVDO_MIN(i + w_ + w_ - 1, n); li1 = li1 + 1; lj1 = i + 1;
        EMIT(1, li1, li2, lj1, lj2);
        if (w_ < 2) continue;
        size_t nbtx = (n - lj2) / (w_ - 1);
        for (size_t k = 0; k < nbtx + 1; ++k) {
            size_t ei = SVDO_MIN(li2 + w_ - 1, m), sj = SVDO_MIN(lj1 + w_ - 1, n), ej3 = SVDO_MIN(lj2 + w_ - 1, n);
            size_t ri1 = li1, ri2 = ei, rj1 = sj, rj2 = lj2;
            li1 = li2; li2 = ei; lj1 = sj; lj2 = ej3;
            if (rj2 > rj1) EMIT(0, ri1, ri2, rj1, rj2);
            if (lj2 > lj1) EMIT(1, li1, li2, lj1, lj2);
        }
    }
#undef EMIT
    return cnt;
}

/* ---------------------------------------------------------------------------------------------
 * QR diagonalisation: svd_serial.h:278-297 (rotate), 314-333 (impl_zero_shift), 138-166
 * (Criteria), 368-422 (qrd).  The reference only compiles for T=float (Rotation is float);
 * the double instantiation here is the same text with T=double and is NOT pinned by the reference.
 * ------------------------------------------------------------------------------------------- */
typedef struct { T c, s, r; } FN(rot);
static FN(rot) FN(rotate)(T u1, T u2) {
    FN(rot) p; T t1, t2, t3;
    if (u1 == 0) { p.c = (T)0.0; p.s = (T)1.0; p.r = u2; }
    else if (FABS_T(u1) > FABS_T(u2)) { t1 = u2 / u1; t2 = SQRT_T(1 + t1 * t1); t3 = 1 / t2; p.c = t3; p.s = t1 * t3; p.r = u1 * t2; }
    else { t1 = u1 / u2; t2 = SQRT_T(1 + t1 * t1); t3 = 1 / t2; p.c = t1 * t3; p.s = t3; p.r = u2 * t2; }
    return p;
}
void FN(svdo_zero_shift)(T* d, T* e, size_t nd) {
    FN(rot) rot = {1, 0, 0}, rot_ = {1, 0, 0};
    for (size_t k = 0; k + 1 < nd; ++k) {
        rot = FN(rotate)(rot.c * d[k], e[k]);
        if (k > 0) e[k - 1] = rot.r * rot_.s;
        rot_ = FN(rotate)(rot_.c * rot.r, d[k + 1] * rot.s);
        d[k] = rot_.r;
    }
    T h = rot.c * d[nd - 1];
    e[nd - 2] = h * rot_.s;
    d[nd - 1] = h * rot_.c;
}
static int FN(cmp_desc)(const void* a, const void* b) {
    T x = *(const T*)a, y = *(const T*)b;
    return (x < y) - (x > y);
}
/* Returns the number of zero-shift sweeps performed (>= 0), or -1 if max_iter was reached
 * (the reference prints an error and returns B unsorted). threshold_out/max_iter_out optional. */
long long FN(svdo_qrd)(T* d, T* e, size_t n, T* threshold_out, unsigned long long* max_iter_out) {
    T eps = (T)1e-8, umin = (T)1e-10, tolerance = 100 * eps;
    T* lambda = (T*)calloc(2 * n, sizeof(T)); T* mu = lambda + n;
    lambda[n - 1] = FABS_T(d[n - 1]);
    for (size_t j = n - 1; j--;) lambda[j] = FABS_T(d[j]) * lambda[j + 1] / (lambda[j + 1] + FABS_T(e[j]));
    mu[0] = FABS_T(d[0]);
    for (size_t j = 0; j + 1 < n; ++j) mu[j + 1] = FABS_T(d[j + 1]) * mu[j] / (mu[j] + FABS_T(e[j]));
    T lmin = lambda[0], mmin = mu[0];
    for (size_t j = 1; j < n; ++j) { if (lambda[j] < lmin) lmin = lambda[j]; if (mu[j] < mmin) mmin = mu[j]; }
    T lbound = lmin < mmin ? lmin : mmin;
    free(lambda);
    size_t max_iter = (500 * n) ^ 2;                   /* svd_serial.h:164: '^' is XOR */
    T a = tolerance * lbound, b2 = (T)max_iter * umin;
    T thr = a < b2 ? b2 : a;                            /* std::max(a,b) */
    if (threshold_out) *threshold_out = thr;
    if (max_iter_out) *max_iter_out = max_iter;

    size_t i_up = n - 2, i_low = 0, j;
    long long sweeps = 0;
    for (size_t iter = 0; iter < max_iter; ++iter) {
        for (size_t i = i_up; i >= 1; --i) { i_up = i; if (FABS_T(e[i]) > thr) break; }
        j = i_up;
        for (size_t i = i_low; i < i_up; ++i) if (FABS_T(e[i]) > thr) { j = i; break; }
        i_low = j;
        if ((i_up == i_low && FABS_T(e[i_up]) <= thr) || (i_up < i_low)) {
            for (size_t q = 0; q < n; ++q) d[q] = FABS_T(d[q]);
            qsort(d, n, sizeof(T), FN(cmp_desc));
            return sweeps;
        }
        FN(svdo_zero_shift)(d + i_low, e + i_low, i_up + 1 - i_low + 1);
        ++sweeps;
    }
    return -1;
}
