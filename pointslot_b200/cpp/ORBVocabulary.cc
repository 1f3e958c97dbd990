// ORBVocabulary.cc — see ORBVocabulary.h.  Host code only; the tree walk runs in libpslot_orb.so (csrc/bow.cu).
#include "ORBVocabulary.h"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>

#include "../../include/pslot_orb.h"

namespace ORB_SLAM2
{

ORBVocabulary::~ORBVocabulary() { pslot_vocab_destroy(h_); }

bool ORBVocabulary::loadFromTextFile(const std::string &filename) // (TemplatedVocabulary.h:1389-1474)
{
    std::ifstream f(filename.c_str());
    if (!f.is_open()) return false;
    std::string s;
    std::getline(f, s);
    int k = 0, L = 0, n1 = -1, n2 = -1;
    {
        std::stringstream ss(s);
        ss >> k >> L >> n1 >> n2;
    }
    if (k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) { // (:1410-1414)
        std::fprintf(stderr, "Vocabulary loading failure: This is not a correct text file!\n");
        return false;
    }
    std::vector<int32_t> parent(1, -1), leaf(1, 0);
    std::vector<unsigned char> desc(32, 0);
    std::vector<double> weight(1, 0.0);
    while (std::getline(f, s)) {
        std::stringstream ss(s);
        int pid, is_leaf;
        if (!(ss >> pid >> is_leaf)) continue; // an empty trailing line: the reference would add a node with an uninitialised parent
        if (pid < 0 || pid >= (int)parent.size()) return false;
        parent.push_back(pid);
        leaf.push_back(is_leaf > 0);
        for (int i = 0; i < 32; ++i) {
            int v = 0;
            ss >> v;
            desc.push_back((unsigned char)v);
        }
        double w = 0.0;
        ss >> w;
        weight.push_back(w);
    }
    const int n = (int)parent.size();
    if (n < 2) return false;
    // children grouped by parent in file order (m_nodes[pid].children.push_back(nid)); leaves numbered in file order
    std::vector<int32_t> start(n + 1, 0), child(n - 1), word(n, -1);
    for (int i = 1; i < n; ++i) ++start[parent[i] + 1];
    for (int i = 0; i < n; ++i) start[i + 1] += start[i];
    {
        std::vector<int32_t> fill(start.begin(), start.end() - 1);
        for (int i = 1; i < n; ++i) child[fill[parent[i]]++] = i;
    }
    unsigned int words = 0;
    for (int i = 1; i < n; ++i)
        if (leaf[i]) word[i] = (int32_t)words++;
    pslot_vocab_destroy(h_);
    h_ = nullptr;
    words_ = 0;
    if (pslot_vocab_create(&h_, n, k, L, n1, n2, desc.data(), start.data(), child.data(), word.data(), weight.data(), 0) != PSLOT_OK) {
        std::fprintf(stderr, "pslot ORBVocabulary: %s\n", pslot_last_error());
        return false;
    }
    words_ = words;
    return true;
}

void ORBVocabulary::transform(const unsigned char *desc, int n, pslot_orb *extractor, DBoW2::BowVector &v,
                              DBoW2::FeatureVector &fv, int levelsup) const
{
    v.clear();
    fv.clear();
    if (empty() || n <= 0) return; // (:1140-1143)
    std::vector<int32_t> leaf(n), nid(n), bid(n), fnode(n), fstart(n + 1), ffeat(n);
    std::vector<double> bval(n);
    int nw = 0, nf = 0;
    if (pslot_bow_transform(h_, extractor, desc, n, levelsup, leaf.data(), nid.data()) != PSLOT_OK ||
        pslot_bow_vectors(h_, leaf.data(), nid.data(), n, bid.data(), bval.data(), &nw, fnode.data(), fstart.data(), ffeat.data(), &nf) != PSLOT_OK) {
        std::fprintf(stderr, "pslot ORBVocabulary::transform: %s\n", pslot_last_error());
        std::abort(); // no error channel on this path, no CPU fallback
    }
    for (int i = 0; i < nw; ++i) v.insert(v.end(), std::make_pair((DBoW2::WordId)bid[i], bval[i]));
    for (int j = 0; j < nf; ++j) {
        std::vector<unsigned int> &dst = fv.insert(fv.end(), std::make_pair((DBoW2::NodeId)fnode[j], std::vector<unsigned int>()))->second;
        dst.assign(ffeat.begin() + fstart[j], ffeat.begin() + fstart[j + 1]);
    }
}

void ORBVocabulary::transform(const std::vector<cv::Mat> &features, DBoW2::BowVector &v, DBoW2::FeatureVector &fv, int levelsup) const
{
    std::vector<unsigned char> flat(32 * features.size());
    for (size_t i = 0; i < features.size(); ++i) std::memcpy(&flat[32 * i], features[i].ptr(0), 32);
    transform(flat.data(), (int)features.size(), nullptr, v, fv, levelsup);
}

} // namespace ORB_SLAM2
