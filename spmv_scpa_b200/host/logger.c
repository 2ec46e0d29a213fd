/* logger.c -- the three CSV sinks.
 *
 * One table describes each file (name, header); one function writes a row.  What reaches the
 * disk is byte-identical to the reference (src/logger.c:31-41 headers, :89-153 rows): CSR rows
 * leave the num_blocks column empty, floats use "%f", files are opened in append mode and get
 * their header only when they did not exist (reference :19-54).
 */
#include <stdarg.h>
#include <stdio.h>
#include <sys/stat.h>

#include "logger.h"

enum sink_id { SINK_SERIAL, SINK_OMP, SINK_CUDA, SINK_COUNT };

static const struct sink_desc {
    const char *file, *label, *header;
} k_sink[SINK_COUNT] = {
    {"serial.csv", "Serial", "matrix,format,rows,cols,nnz,num_blocks,duration_ms,gflops"},
    {"omp.csv", "OMP", "matrix,format,bench,rows,cols,nnz,num_blocks,num_threads,duration_ms,gflops"},
    {"cuda.csv", "CUDA",
     "matrix,format,kernel,warps_per_block,rows,cols,nnz,num_blocks,duration_ms,gflops"},
};

static FILE *g_file[SINK_COUNT];

int logger_init(const char *directory) {
    int missing = 0;
    for (int s = 0; s < SINK_COUNT; ++s) {
        char path[MAX_PATH];
        struct stat st;
        snprintf(path, sizeof path, "%s/%s", directory, k_sink[s].file);
        const int existed = stat(path, &st) == 0;
        g_file[s] = fopen(path, "a");
        if (g_file[s] == NULL) {
            ++missing;
            continue;
        }
        if (!existed) {
            fprintf(g_file[s], "%s\n", k_sink[s].header);
            fflush(g_file[s]);
        }
    }
    return missing ? -1 : 0;
}

void logger_close(void) {
    for (int s = 0; s < SINK_COUNT; ++s) {
        if (g_file[s] != NULL)
            fclose(g_file[s]);
        g_file[s] = NULL;
    }
}

/* append one formatted row to a sink and flush it */
static void row(enum sink_id s, const char *fmt, ...) {
    if (g_file[s] == NULL) {
        LOG_ERR("%s log not initialized", k_sink[s].label);
        return;
    }
    va_list ap;
    va_start(ap, fmt);
    vfprintf(g_file[s], fmt, ap);
    va_end(ap);
    fflush(g_file[s]);
}

/* the CSR rows have no hack count: the num_blocks column stays empty (",,") */

void log_csr_serial_benchmark(const sparse_csr *m, bench r) {
    row(SINK_SERIAL, "%s,CSR,%d,%d,%d,,%f,%f\n", m->name, m->M, m->N, m->NZ, r.duration_ms, r.gflops);
}

void log_hll_serial_benchmark(const sparse_hll *m, bench r) {
    row(SINK_SERIAL, "%s,HLL,%d,%d,%d,%d,%f,%f\n", m->name, m->M, m->N, m->NZ, m->num_blocks,
        r.duration_ms, r.gflops);
}

void log_csr_omp_benchmark(const sparse_csr *m, bench_omp r) {
    row(SINK_OMP, "%s,CSR,%s,%d,%d,%d,,%d,%f,%f\n", m->name, r.name, m->M, m->N, m->NZ,
        r.num_threads, r.bench.duration_ms, r.bench.gflops);
}

void log_hll_omp_benchmark(const sparse_hll *m, bench_omp r) {
    row(SINK_OMP, "%s,HLL,%s,%d,%d,%d,%d,%d,%f,%f\n", m->name, r.name, m->M, m->N, m->NZ,
        m->num_blocks, r.num_threads, r.bench.duration_ms, r.bench.gflops);
}

void log_csr_cuda_benchmark(const sparse_csr *m, bench_cuda r, int kernel_id) {
    row(SINK_CUDA, "%s,CSR,%d,%d,%d,%d,%d,,%f,%f\n", m->name, kernel_id, r.warps_per_block, m->M,
        m->N, m->NZ, r.bench.duration_ms, r.bench.gflops);
}

void log_hll_cuda_benchmark(const sparse_hll *m, bench_cuda r, int kernel_id) {
    row(SINK_CUDA, "%s,HLL,%d,%d,%d,%d,%d,%d,%f,%f\n", m->name, kernel_id, r.warps_per_block, m->M,
        m->N, m->NZ, m->num_blocks, r.bench.duration_ms, r.bench.gflops);
}
