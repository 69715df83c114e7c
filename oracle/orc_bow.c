/*
 * orc_bow.c -- oracle restatement of DBoW2's vocabulary-tree descent and bag-of-words maps:
 * TemplatedVocabulary::transform(feature, word_id, weight, nid, levelsup)
 * (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1218-1258), the feature loop of
 * transform(features, BowVector&, FeatureVector&, levelsup) (:1124-1190) for TF_IDF / TF weighting with L1
 * normalisation (BowVector.cpp:34-46, :62-84) and FeatureVector::addFeature (FeatureVector.cpp:31-45).
 * TEST INFRASTRUCTURE ONLY (see plf_oracle.h).  Pinned to our reading of the source (no DBoW2 build exists here).
 * Tree layout = what loadFromTextFile (:1338-1424) builds: children in node-id order, words in leaf order.
 */
#include "plf_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

void orc_bow_transform(int L, int nnodes, const int32_t* parent, const uint8_t* ndesc, const double* nweight, const uint8_t* is_leaf,
                       const uint8_t* feat, int n, int levelsup, int32_t* word, double* weight, int32_t* node)
{
    /* children lists (push_back order) and word ids */
    int* off = (int*)calloc((size_t)nnodes + 2, sizeof(int));
    int* ids = (int*)malloc(sizeof(int) * (size_t)nnodes);
    int* fill = (int*)calloc((size_t)nnodes + 1, sizeof(int));
    int* wid = (int*)malloc(sizeof(int) * (size_t)nnodes);
    for (int i = 1; i < nnodes; i++) off[parent[i] + 1]++;
    for (int i = 0; i < nnodes; i++) off[i + 1] += off[i];
    for (int i = 1; i < nnodes; i++) ids[off[parent[i]] + fill[parent[i]]++] = i;
    int nw = 0;
    for (int i = 0; i < nnodes; i++) wid[i] = (i > 0 && is_leaf[i]) ? nw++ : -1;
    const int nid_level = L - levelsup;
    for (int f = 0; f < n; f++) {
        const uint8_t* a = feat + (size_t)f * 32;
        int final_id = 0, current_level = 0, nid = 0;
        while (off[final_id + 1] > off[final_id]) {          /* do { } while (!isLeaf()) from the root */
            ++current_level;
            const int c0 = off[final_id], c1 = off[final_id + 1];
            final_id = ids[c0];
            double best_d = (double)orc_descriptor_distance(a, ndesc + (size_t)final_id * 32);
            for (int c = c0 + 1; c < c1; c++) {
                const int id = ids[c];
                const double d = (double)orc_descriptor_distance(a, ndesc + (size_t)id * 32);
                if (d < best_d) { best_d = d; final_id = id; }
            }
            if (current_level == nid_level) nid = final_id;
        }
        word[f] = wid[final_id];
        weight[f] = nweight[final_id];
        node[f] = nid;
    }
    free(off); free(ids); free(fill); free(wid);
}
