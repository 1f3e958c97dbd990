#!/usr/bin/env python3
"""Lift whole function definitions out of the REFERENCE tree at build time (test infrastructure).

`Frame.cc` and `ORBmatcher.cc` cannot be compiled as files (they include g2o, DBoW2, TensorRT, opencv contrib ...), but
the handful of member functions on the hot path only touch plain members.  This script copies those function bodies
VERBATIM - located by their signature line, ended by brace matching - from where they lie under the reference tree
into a scratch directory; `ref_frame_glue.cc` #includes them inside stub classes that carry exactly the members they
read (frame_stub.h).  Nothing extracted is ever written into the repository: the Makefile points `--out` at a mktemp
directory and removes it after the compile.

usage: extract_fn.py --ref /root/reference --out DIR
"""
import argparse
import re
import sys
from pathlib import Path

# (output file, source file relative to the reference root, [signature prefixes, matched at column 0])
PIECES = [
    ("frame_fns.inc", "src/Frame.cc", [
        "void Frame::ComputeStereoMatches()",
        "void Frame::ComputeObjStereoMatches()",
        "void Frame::AssignFeaturesToGrid()",
        "vector<size_t> Frame::GetFeaturesInArea(const float &x, const float  &y, const float  &r, const int minLevel, const int maxLevel) const",
        "bool Frame::PosInGrid(const cv::KeyPoint &kp, int &posX, int &posY)",
        "vector<size_t> Frame::GetObjectFeaturesInArea(const int& nDetObjOrder, const size_t& nDetObjFeaNum, const float &x, const float  &y, const float  &r, const int minLevel, const int maxLevel) const",
        "bool Frame::isInBBox(const size_t &nOrder, const float &u, const float &v)",
        "void Frame::AssignFeatures(vector<DetectionObject*> &vDetectionObjects)",
        "void Frame::AssignDetObjFeasToGrid(const size_t &objOrder, const size_t &objFeaNum)",
    ]),
    ("keyframe_fns.inc", "src/KeyFrame.cc", [
        "vector<size_t> KeyFrame::GetFeaturesInArea(const float &x, const float &y, const float &r) const",
        "bool KeyFrame::IsInImage(const float &x, const float &y) const",
    ]),
    ("matcher_fns.inc", "src/ORBmatcher.cc", [
        "int ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, const float th)",
        "int ORBmatcher::SearchByProjection(Frame &F, const size_t &nOrder, const vector<MapObjectPoint*> &vpMapPoints, const float th)",
        "int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)",
        "int ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const set<MapPoint*> &sAlreadyFound, const float th , const int ORBdist)",
        "int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const vector<MapPoint*> &vpPoints, vector<MapPoint*> &vpMatched, int th)",
        "float ORBmatcher::RadiusByViewingCos(const float &viewCos)",
        "void ORBmatcher::ComputeThreeMaxima(vector<int>* histo, const int L, int &ind1, int &ind2, int &ind3)",
        "int ORBmatcher::DescriptorDistance(const cv::Mat &a, const cv::Mat &b)",
    ]),
]
# one-line constant definitions (ORBmatcher.cc:56-60)
CONSTS = ("matcher_consts.inc", "src/ORBmatcher.cc", re.compile(r"^const int ORBmatcher::\w+ = \d+;\s*$"))


def extract(text: str, sig: str) -> str:
    lines = text.split("\n")
    hits = [i for i, l in enumerate(lines) if l.rstrip() == sig]
    if len(hits) != 1:
        raise SystemExit(f"extract_fn: signature found {len(hits)} times: {sig}")
    i = hits[0]
    depth, started, out = 0, False, []
    for l in lines[i:]:
        out.append(l)
        # the hot-path functions hold no braces inside string/char literals or comments; checked by the balance below
        depth += l.count("{") - l.count("}")
        started = started or "{" in l
        if started and depth == 0:
            return "\n".join(out) + "\n"
    raise SystemExit(f"extract_fn: unbalanced braces after: {sig}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", required=True)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    ref, out = Path(a.ref), Path(a.out)
    out.mkdir(parents=True, exist_ok=True)
    for name, src, sigs in PIECES:
        text = (ref / src).read_text(encoding="utf-8", errors="replace")
        body = "".join(f"// ---- {src}: {s}\n" + extract(text, s) + "\n" for s in sigs)
        (out / name).write_text(body, encoding="utf-8")
    name, src, rx = CONSTS
    text = (ref / src).read_text(encoding="utf-8", errors="replace")
    consts = [l for l in text.split("\n") if rx.match(l)]
    if len(consts) < 3:
        raise SystemExit("extract_fn: ORBmatcher constants not found")
    (out / name).write_text("\n".join(consts) + "\n", encoding="utf-8")
    return 0


if __name__ == "__main__":
    sys.exit(main())
