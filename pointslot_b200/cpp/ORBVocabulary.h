// ORBVocabulary.h — the vocabulary side of Frame::ComputeBoW (reference include/ORBVocabulary.h:30:
// `typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary`; members used on this path:
// loadFromTextFile (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1389-1474, called from src/System.cc) and
// transform(features, BowVector&, FeatureVector&, levelsup) (:1132-1201)) as a drop-in over pslot_vocab_* /
// pslot_bow_*: same names, same signatures, same results (word values bit for bit).  BowVector / FeatureVector keep
// DBoW2's definitions (std::map, BowVector.h:78, FeatureVector.h:24); in a PointSLOT tree the real DBoW2 headers provide
// them and only the two member bodies change.
#ifndef ORBVOCABULARY_H
#define ORBVOCABULARY_H

#include <map>
#include <string>
#include <vector>

#include "compat/opencv_min.h"

struct pslot_vocab;
struct pslot_orb;

namespace DBoW2
{
typedef unsigned int WordId;
typedef double WordValue;
typedef unsigned int NodeId;
class BowVector : public std::map<WordId, WordValue> {};
class FeatureVector : public std::map<NodeId, std::vector<unsigned int>> {};
} // namespace DBoW2

namespace ORB_SLAM2
{

class ORBVocabulary
{
public:
    ORBVocabulary() {}
    ~ORBVocabulary();
    ORBVocabulary(const ORBVocabulary &) = delete;
    ORBVocabulary &operator=(const ORBVocabulary &) = delete;

    // `k L scoring weighting`, then one `parent isLeaf d0 .. d31 weight` line per node; the tree goes to the device
    bool loadFromTextFile(const std::string &filename);
    bool empty() const { return words_ == 0; }
    unsigned int size() const { return words_; }

    // DBoW2's signature: one 1 x 32 CV_8U row per feature
    void transform(const std::vector<cv::Mat> &features, DBoW2::BowVector &v, DBoW2::FeatureVector &fv, int levelsup) const;
    // the same on n descriptors (row-major, 32 bytes each) in host memory, or - desc == nullptr - on the descriptors the
    // extractor's last operator() left on the device (what Frame::ComputeBoW uses: they never travel)
    void transform(const unsigned char *desc, int n, pslot_orb *extractor, DBoW2::BowVector &v, DBoW2::FeatureVector &fv,
                   int levelsup) const;

private:
    pslot_vocab *h_ = nullptr;
    unsigned int words_ = 0;
};

} // namespace ORB_SLAM2
#endif
