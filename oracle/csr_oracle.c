/*
 * CPU oracle (C restatement) of the destination-sorted CSR / source-sorted CSC construction.
 *
 * TEST INFRASTRUCTURE ONLY: linked by tests/ and __graft_entry__.build(); never by the product path.
 *
 * Restates what PyG's GATConv.forward does to edge_index on every call at the reference's call sites
 * (src/models/gat.py:80, src/models/tgn.py:94): remove_self_loops (order-preserving) followed by
 * add_self_loops (arange(N) appended last), and then the ordering torch.sort(dst', stable=True) would
 * give.  A stable counting sort is by construction the same permutation as a stable comparison sort,
 * so this is an independent second statement of oracle/pyg_gatconv.py::csr_oracle / csc_oracle.
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/libgnnfd_csr_oracle.so oracle/csr_oracle.c
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Returns E' (number of edges after the rewrite) or -1 on a bad index.
 * rowptr[N+1], col[cap], perm[cap] with cap >= E + N. */
int64_t gnnfd_oracle_csr(const int64_t* edge_index, int64_t E, int64_t N, int add_self_loops,
                         int64_t* rowptr, int64_t* col, int64_t* perm)
{
    const int64_t* src = edge_index;
    const int64_t* dst = edge_index + E;
    int64_t cap = E + (add_self_loops ? N : 0);
    int64_t* s2 = (int64_t*)malloc(sizeof(int64_t) * (size_t)(cap > 0 ? cap : 1));
    int64_t* d2 = (int64_t*)malloc(sizeof(int64_t) * (size_t)(cap > 0 ? cap : 1));
    int64_t Ep = 0;
    for (int64_t e = 0; e < E; ++e) {
        if (src[e] < 0 || src[e] >= N || dst[e] < 0 || dst[e] >= N) { free(s2); free(d2); return -1; }
        if (add_self_loops && src[e] == dst[e]) continue;      /* remove_self_loops */
        s2[Ep] = src[e]; d2[Ep] = dst[e]; ++Ep;
    }
    if (add_self_loops)
        for (int64_t n = 0; n < N; ++n) { s2[Ep] = n; d2[Ep] = n; ++Ep; }   /* add_self_loops */

    memset(rowptr, 0, sizeof(int64_t) * (size_t)(N + 1));
    for (int64_t e = 0; e < Ep; ++e) rowptr[d2[e] + 1]++;
    for (int64_t n = 0; n < N; ++n) rowptr[n + 1] += rowptr[n];
    int64_t* cursor = (int64_t*)malloc(sizeof(int64_t) * (size_t)(N > 0 ? N : 1));
    memcpy(cursor, rowptr, sizeof(int64_t) * (size_t)N);
    for (int64_t e = 0; e < Ep; ++e) {                          /* stable: ascending e within a row */
        int64_t p = cursor[d2[e]]++;
        perm[p] = e; col[p] = s2[e];
    }
    free(cursor); free(s2); free(d2);
    return Ep;
}

/* Source-major twin: stable sort of the CSR-ordered col array.  colptr[N+1], row[Ep], eid[Ep]. */
void gnnfd_oracle_csc(const int64_t* rowptr, const int64_t* col, int64_t N, int64_t Ep,
                      int64_t* colptr, int64_t* row, int64_t* eid)
{
    memset(colptr, 0, sizeof(int64_t) * (size_t)(N + 1));
    for (int64_t e = 0; e < Ep; ++e) colptr[col[e] + 1]++;
    for (int64_t n = 0; n < N; ++n) colptr[n + 1] += colptr[n];
    int64_t* cursor = (int64_t*)malloc(sizeof(int64_t) * (size_t)(N > 0 ? N : 1));
    memcpy(cursor, colptr, sizeof(int64_t) * (size_t)N);
    for (int64_t i = 0; i < N; ++i)
        for (int64_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
            int64_t p = cursor[col[e]]++;
            eid[p] = e; row[p] = i;
        }
    free(cursor);
}
