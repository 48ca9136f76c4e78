/*
 * oracle/par_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of gzp's ParCompress thread topology
 * (/root/reference/src/par/compress.rs): the caller thread chunks the input
 * (`write`, :413-463, strict '>' hold-back; `flush_last(true)`, :332-362),
 * `num_threads` workers run FormatSpec::encode + Check::update (:267-300), one
 * writer drains a FIFO of tickets in submission order and folds the running
 * check (:303-313).  Both channels are bounded at 2*num_threads (:111-112).
 * Used as the CPU baseline ("port") by bench.py; never by the product path.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "oracle.h"

#define DICT_SIZE 32768

typedef struct {
    const uint8_t *buf; size_t len; const uint8_t *dict; size_t dict_len; int is_last;
    uint8_t *out; long out_len; uint32_t sum, amount; int done;
} ticket_t;

typedef struct {
    pthread_mutex_t mu; pthread_cond_t not_empty, not_full;
    ticket_t **ring; size_t cap, head, tail; int closed;
} chan_t;

static void chan_init(chan_t *c, size_t cap) { pthread_mutex_init(&c->mu, 0); pthread_cond_init(&c->not_empty, 0); pthread_cond_init(&c->not_full, 0); c->ring = calloc(cap, sizeof(*c->ring)); c->cap = cap; c->head = c->tail = 0; c->closed = 0; }
static void chan_send(chan_t *c, ticket_t *t)
{
    pthread_mutex_lock(&c->mu);
    while (c->tail - c->head == c->cap) pthread_cond_wait(&c->not_full, &c->mu);
    c->ring[c->tail++ % c->cap] = t;
    pthread_cond_signal(&c->not_empty);
    pthread_mutex_unlock(&c->mu);
}
static ticket_t *chan_recv(chan_t *c)
{
    ticket_t *t = 0;
    pthread_mutex_lock(&c->mu);
    while (c->tail == c->head && !c->closed) pthread_cond_wait(&c->not_empty, &c->mu);
    if (c->tail != c->head) { t = c->ring[c->head++ % c->cap]; pthread_cond_signal(&c->not_full); }
    pthread_mutex_unlock(&c->mu);
    return t;
}
static void chan_close(chan_t *c) { pthread_mutex_lock(&c->mu); c->closed = 1; pthread_cond_broadcast(&c->not_empty); pthread_mutex_unlock(&c->mu); }

typedef struct {
    int format, level; chan_t work, order;
    pthread_mutex_t dmu; pthread_cond_t dcv;
    uint8_t *out; size_t out_cap, out_len; int err;
} par_t;

static void *worker(void *arg)
{
    par_t *p = arg; ticket_t *t;
    while ((t = chan_recv(&p->work))) {
        size_t cap = oracle_encode_capacity(p->format, t->len);
        t->out = malloc(cap ? cap : 1);
        t->out_len = oracle_encode_block(p->format, p->level, t->buf, t->len, t->dict, t->dict_len, t->is_last, t->out, cap);
        if (p->format == ORACLE_FMT_GZIP) { t->sum = oracle_crc32(0, t->buf, t->len); t->amount = (uint32_t)t->len; }
        else if (p->format == ORACLE_FMT_ZLIB) { t->sum = oracle_adler32(1, t->buf, t->len); t->amount = (uint32_t)t->len; }
        pthread_mutex_lock(&p->dmu); t->done = 1; pthread_cond_broadcast(&p->dcv); pthread_mutex_unlock(&p->dmu);
    }
    oracle_thread_cleanup();
    return 0;
}

static void *writer(void *arg)
{
    par_t *p = arg; ticket_t *t;
    uint32_t sum = p->format == ORACLE_FMT_ZLIB ? 1 : 0, amount = 0;
    uint8_t hb[16];
    size_t h = oracle_header(p->format, p->level, hb);
    if (p->out_len + h <= p->out_cap) memcpy(p->out + p->out_len, hb, h); else p->err = 1;
    p->out_len += h;
    while ((t = chan_recv(&p->order))) {
        pthread_mutex_lock(&p->dmu);
        while (!t->done) pthread_cond_wait(&p->dcv, &p->dmu);
        pthread_mutex_unlock(&p->dmu);
        if (t->out_len < 0) p->err = (int)t->out_len;
        else {
            if (p->format == ORACLE_FMT_GZIP) sum = oracle_crc32_combine(sum, t->sum, t->amount);
            else if (p->format == ORACLE_FMT_ZLIB) sum = oracle_adler32_combine(sum, t->sum, t->amount);
            amount += t->amount;
            if (p->out_len + (size_t)t->out_len <= p->out_cap) memcpy(p->out + p->out_len, t->out, (size_t)t->out_len); else p->err = 1;
            p->out_len += (size_t)t->out_len;
        }
        free(t->out); free(t);
    }
    h = oracle_footer(p->format, sum, amount, hb);
    if (p->out_len + h <= p->out_cap) memcpy(p->out + p->out_len, hb, h); else p->err = 1;
    p->out_len += h;
    return 0;
}

static int needs_dict(int f) { return f == ORACLE_FMT_GZIP || f == ORACLE_FMT_ZLIB || f == ORACLE_FMT_RAWDEFLATE; }

/* One whole-stream pass: write(in) then finish().  Returns seconds (wall), <0 on error. */
double oracle_par_compress(int format, int level, size_t buffer_size, int num_threads, const uint8_t *in,
                           size_t n, uint8_t *out, size_t out_cap, size_t *out_len)
{
    par_t p; memset(&p, 0, sizeof p);
    p.format = format; p.level = level; p.out = out; p.out_cap = out_cap;
    if (num_threads < 1) num_threads = 1;
    chan_init(&p.work, (size_t)num_threads * 2); chan_init(&p.order, (size_t)num_threads * 2);
    pthread_mutex_init(&p.dmu, 0); pthread_cond_init(&p.dcv, 0);
    struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_t *th = calloc((size_t)num_threads, sizeof *th), wr;
    for (int i = 0; i < num_threads; i++) pthread_create(&th[i], 0, worker, &p);
    pthread_create(&wr, 0, writer, &p);

    size_t pos = 0; const uint8_t *dict = 0; size_t dict_len = 0;
    /* write(): emit full blocks while MORE than buffer_size bytes are buffered */
    while (n - pos > buffer_size) {
        ticket_t *t = calloc(1, sizeof *t);
        t->buf = in + pos; t->len = buffer_size; t->dict = dict; t->dict_len = dict_len;
        if (needs_dict(format)) { dict = t->buf + t->len - DICT_SIZE; dict_len = DICT_SIZE; } else { dict = 0; dict_len = 0; }
        chan_send(&p.order, t); chan_send(&p.work, t);
        pos += buffer_size;
    }
    /* finish(): flush_last(true) — always at least one (possibly empty) block */
    for (;;) {
        ticket_t *t = calloc(1, sizeof *t);
        size_t len = n - pos < buffer_size ? n - pos : buffer_size;
        t->buf = in + pos; t->len = len; t->dict = dict; t->dict_len = dict_len; dict = 0; dict_len = 0;
        pos += len;
        if (pos == n) t->is_last = 1;
        if (t->len >= DICT_SIZE && !t->is_last && needs_dict(format)) { dict = t->buf + t->len - DICT_SIZE; dict_len = DICT_SIZE; }
        chan_send(&p.order, t); chan_send(&p.work, t);
        if (pos == n) break;
    }
    chan_close(&p.work); chan_close(&p.order);
    for (int i = 0; i < num_threads; i++) pthread_join(th[i], 0);
    pthread_join(wr, 0);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(th); free(p.work.ring); free(p.order.ring);
    if (out_len) *out_len = p.out_len;
    if (p.err) return -1.0;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
