// ref_dbow_glue.cc — C entry points over the REFERENCE's own vendored DBoW2 (Thirdparty/DBoW2, compiled from where it lies
// under $(REF) by `make -C oracle ref`; test infrastructure, never on the product path): the vocabulary PointSLOT uses
// (include/ORBVocabulary.h: TemplatedVocabulary<FORB::TDescriptor, FORB>), loaded with loadFromTextFile as System.cc does,
// and Frame::ComputeBoW's call (src/Frame.cc:2040-2047): transform(toDescriptorVector(mDescriptors), mBowVec, mFeatVec, 4).
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "Thirdparty/DBoW2/DBoW2/FORB.h"
#include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary; // reference include/ORBVocabulary.h:30

extern "C" {

void *refbow_load(const char *path)
{
    ORBVocabulary *v = new ORBVocabulary();
    if (!v->loadFromTextFile(path)) { delete v; return nullptr; }
    return v;
}
void refbow_free(void *v) { delete (ORBVocabulary *)v; }
int refbow_size(void *v) { return (int)((ORBVocabulary *)v)->size(); }

// mBowVec as (ids ascending, values), mFeatVec as CSR; arrays sized n (fv_start n + 1).  Returns the number of words.
int refbow_transform(void *voc, const uint8_t *desc, int n, int levelsup, int32_t *bow_id, double *bow_val, int32_t *fv_node,
                     int32_t *fv_start, int32_t *fv_feat, int *n_fv)
{
    std::vector<cv::Mat> vDesc; // Converter::toDescriptorVector (src/Converter.cc): one 1 x 32 row per feature
    vDesc.reserve((size_t)n);
    for (int i = 0; i < n; ++i) {
        cv::Mat m(1, 32, CV_8U);
        std::memcpy(m.ptr(), desc + (size_t)i * 32, 32);
        vDesc.push_back(m);
    }
    DBoW2::BowVector bv;
    DBoW2::FeatureVector fv;
    ((ORBVocabulary *)voc)->transform(vDesc, bv, fv, levelsup);
    int k = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++k) { bow_id[k] = (int32_t)it->first; bow_val[k] = it->second; }
    int j = 0, pos = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++j) {
        fv_node[j] = (int32_t)it->first;
        fv_start[j] = pos;
        for (size_t t = 0; t < it->second.size(); ++t) fv_feat[pos++] = (int32_t)it->second[t];
    }
    fv_start[j] = pos;
    *n_fv = j;
    return k;
}

} // extern "C"
