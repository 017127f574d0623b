/*
 * oracle/shim/sonLib.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Minimal stand-in for the external (un-vendored, unpinned) sonLib container
 * library that the reference links against (reference Dockerfile:21-25).
 * Only the container entry points that the hmm_flagger call graph touches are
 * provided (stList / stHash / stSet).  Written from the published sonLib API;
 * no arithmetic lives here.  Used solely to build the UNMODIFIED reference
 * sources from /root/reference into oracle/_ref/.
 */
#ifndef HFG_ORACLE_SONLIB_SHIM_H
#define HFG_ORACLE_SONLIB_SHIM_H

#include <stdint.h>
#include <stdbool.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct _stList stList;
typedef struct _stHash stHash;
typedef struct _stHashIterator stHashIterator;
typedef struct _stHash stSet;

stList *stList_construct3(int64_t length, void (*destructElement)(void *));
void stList_append(stList *list, void *item);
void *stList_get(stList *list, int64_t index);
void stList_set(stList *list, int64_t index, void *item);
int64_t stList_length(stList *list);
void stList_destruct(stList *list);
void stList_sort(stList *list, int (*cmpFn)(const void *a, const void *b));

stHash *stHash_construct3(uint64_t (*hashKey)(const void *),
                          int (*hashEqualsKey)(const void *, const void *),
                          void (*destructKeys)(void *),
                          void (*destructValues)(void *));
void stHash_insert(stHash *hash, void *key, void *value);
void *stHash_search(stHash *hash, void *key);
void stHash_destruct(stHash *hash);
stHashIterator *stHash_getIterator(stHash *hash);
void *stHash_getNext(stHashIterator *iterator);
void stHash_destructIterator(stHashIterator *iterator);
stList *stHash_getKeys(stHash *hash);
uint64_t stHash_stringKey(const void *k);
int stHash_stringEqualKey(const void *key1, const void *key2);

stSet *stSet_construct3(uint64_t (*hashKey)(const void *),
                        int (*hashEqualsKey)(const void *, const void *),
                        void (*destructKeys)(void *));
void stSet_insert(stSet *set, void *key);
void *stSet_search(stSet *set, void *key);

#ifdef __cplusplus
}
#endif
#endif
