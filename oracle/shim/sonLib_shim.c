/*
 * oracle/shim/sonLib_shim.c -- TEST INFRASTRUCTURE ONLY (see sonLib.h).
 * Growable pointer array + chained hash table; insertion-ordered iteration.
 */
#include "sonLib.h"

struct _stList {
    void **items;
    int64_t n, cap;
    void (*destructElement)(void *);
};

stList *stList_construct3(int64_t length, void (*destructElement)(void *)) {
    stList *l = calloc(1, sizeof(stList));
    l->cap = length > 8 ? length : 8;
    l->items = calloc((size_t) l->cap, sizeof(void *));
    l->n = length; /* "length" NULL slots, as in sonLib */
    l->destructElement = destructElement;
    return l;
}

void stList_append(stList *l, void *item) {
    if (l->n == l->cap) {
        l->cap *= 2;
        l->items = realloc(l->items, (size_t) l->cap * sizeof(void *));
    }
    l->items[l->n++] = item;
}

void *stList_get(stList *l, int64_t i) {
    if (i < 0 || i >= l->n) {
        fprintf(stderr, "[sonLib shim] stList_get index %ld out of range (%ld)\n", (long) i, (long) l->n);
        abort();
    }
    return l->items[i];
}

void stList_set(stList *l, int64_t i, void *item) {
    if (i < 0 || i >= l->n) {
        fprintf(stderr, "[sonLib shim] stList_set index %ld out of range (%ld)\n", (long) i, (long) l->n);
        abort();
    }
    l->items[i] = item;
}

int64_t stList_length(stList *l) { return l == NULL ? 0 : l->n; }

void stList_destruct(stList *l) {
    if (l == NULL) return;
    if (l->destructElement != NULL) {
        for (int64_t i = 0; i < l->n; i++) {
            if (l->items[i] != NULL) l->destructElement(l->items[i]); /* NULL slots skipped */
        }
    }
    free(l->items);
    free(l);
}

/* the comparator receives the ELEMENTS, not pointers to the slots */
void stList_sort(stList *l, int (*cmpFn)(const void *a, const void *b)) {
    /* stable merge sort so equal elements keep their order */
    int64_t n = l->n;
    if (n < 2) return;
    void **tmp = malloc((size_t) n * sizeof(void *));
    for (int64_t w = 1; w < n; w *= 2) {
        for (int64_t lo = 0; lo < n; lo += 2 * w) {
            int64_t mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n;
            int64_t a = lo, b = mid, k = lo;
            while (a < mid && b < hi) tmp[k++] = cmpFn(l->items[b], l->items[a]) < 0 ? l->items[b++] : l->items[a++];
            while (a < mid) tmp[k++] = l->items[a++];
            while (b < hi) tmp[k++] = l->items[b++];
        }
        memcpy(l->items, tmp, (size_t) n * sizeof(void *));
    }
    free(tmp);
}

typedef struct HNode {
    void *key, *value;
    struct HNode *next;      /* bucket chain */
    struct HNode *ordNext;   /* insertion order */
} HNode;

struct _stHash {
    HNode **buckets;
    uint64_t nb, n;
    HNode *first, *last;
    uint64_t (*hashKey)(const void *);
    int (*eq)(const void *, const void *);
    void (*dk)(void *);
    void (*dv)(void *);
};

struct _stHashIterator {
    HNode *cur;
};

static uint64_t ptrKey(const void *k) { return (uint64_t) (uintptr_t) k * 0x9E3779B97F4A7C15ULL; }
static int ptrEq(const void *a, const void *b) { return a == b; }

stHash *stHash_construct3(uint64_t (*hashKey)(const void *), int (*eq)(const void *, const void *),
                          void (*dk)(void *), void (*dv)(void *)) {
    stHash *h = calloc(1, sizeof(stHash));
    h->nb = 1024;
    h->buckets = calloc(h->nb, sizeof(HNode *));
    h->hashKey = hashKey ? hashKey : ptrKey;
    h->eq = eq ? eq : ptrEq;
    h->dk = dk;
    h->dv = dv;
    return h;
}

static void rehash(stHash *h) {
    uint64_t nb = h->nb * 4;
    HNode **b = calloc(nb, sizeof(HNode *));
    for (HNode *x = h->first; x; x = x->ordNext) {
        uint64_t i = h->hashKey(x->key) % nb;
        x->next = b[i];
        b[i] = x;
    }
    free(h->buckets);
    h->buckets = b;
    h->nb = nb;
}

void stHash_insert(stHash *h, void *key, void *value) {
    uint64_t i = h->hashKey(key) % h->nb;
    for (HNode *x = h->buckets[i]; x; x = x->next) {
        if (h->eq(x->key, key)) { /* replace */
            if (h->dv && x->value != value) h->dv(x->value);
            if (h->dk && x->key != key) h->dk(x->key);
            x->key = key;
            x->value = value;
            return;
        }
    }
    HNode *x = calloc(1, sizeof(HNode));
    x->key = key;
    x->value = value;
    x->next = h->buckets[i];
    h->buckets[i] = x;
    if (h->last) h->last->ordNext = x; else h->first = x;
    h->last = x;
    if (++h->n > 2 * h->nb) rehash(h);
}

void *stHash_search(stHash *h, void *key) {
    uint64_t i = h->hashKey(key) % h->nb;
    for (HNode *x = h->buckets[i]; x; x = x->next)
        if (h->eq(x->key, key)) return x->value;
    return NULL;
}

void stHash_destruct(stHash *h) {
    if (h == NULL) return;
    HNode *x = h->first;
    while (x) {
        HNode *nx = x->ordNext;
        if (h->dk && x->key) h->dk(x->key);
        if (h->dv && x->value) h->dv(x->value);
        free(x);
        x = nx;
    }
    free(h->buckets);
    free(h);
}

stHashIterator *stHash_getIterator(stHash *h) {
    stHashIterator *it = calloc(1, sizeof(stHashIterator));
    it->cur = h->first;
    return it;
}

void *stHash_getNext(stHashIterator *it) { /* returns KEYS */
    if (it->cur == NULL) return NULL;
    void *k = it->cur->key;
    it->cur = it->cur->ordNext;
    return k;
}

void stHash_destructIterator(stHashIterator *it) { free(it); }

stList *stHash_getKeys(stHash *h) {
    stList *l = stList_construct3(0, NULL);
    for (HNode *x = h->first; x; x = x->ordNext) stList_append(l, x->key);
    return l;
}

uint64_t stHash_stringKey(const void *k) {
    const unsigned char *s = k;
    uint64_t hsh = 1469598103934665603ULL;
    while (*s) { hsh ^= *s++; hsh *= 1099511628211ULL; }
    return hsh;
}

int stHash_stringEqualKey(const void *a, const void *b) { return strcmp(a, b) == 0; }

stSet *stSet_construct3(uint64_t (*hashKey)(const void *), int (*eq)(const void *, const void *),
                        void (*dk)(void *)) {
    return stHash_construct3(hashKey, eq, dk, NULL);
}
void stSet_insert(stSet *s, void *key) { stHash_insert(s, key, key); }
void *stSet_search(stSet *s, void *key) { return stHash_search(s, key); }
