// Frame.h — the stereo-construction path of PointSLOT's Frame (reference include/Frame.h:65, 94, 134,
// 232-273, 303-325; src/Frame.cc:144-200, 1658-1664, 2049-2079, 2142-2316) as a drop-in:
// same constructor signature, same member names, same outputs (mvKeys, mvKeysRight, mvKeysUn, mDescriptors,
// mDescriptorsRight, mvuRight, mvDepth, N, mbf, mb).  Everything else in the reference's Frame (object
// modes, grid, BoW, pose) consumes these members and is out of scope; in a PointSLOT tree only the three
// member functions below are replaced.
#ifndef FRAME_H
#define FRAME_H

#include <vector>

#include "ORBextractor.h"
#include "compat/opencv_min.h"

namespace ORB_SLAM2
{

class ORBVocabulary; // opaque here: the stereo path never touches it

class Frame
{
public:
    Frame();

    // Constructor for stereo cameras (reference Frame.h:65 / Frame.cc:144-200).
    Frame(const cv::Mat &imLeft, const cv::Mat &imRight, const double &timeStamp, ORBextractor *extractorLeft,
          ORBextractor *extractorRight, ORBVocabulary *voc, cv::Mat &K, cv::Mat &distCoef, const float &bf,
          const float &thDepth);

    // Extract ORB on the image. 0 for left image and 1 for right image.
    void ExtractORB(int flag, const cv::Mat &im);

    // Search a match for each keypoint in the left image to a keypoint in the right image.
    // If there is a match, depth is computed and the right coordinate associated to the left keypoint is stored.
    void ComputeStereoMatches();

public:
    ORBVocabulary *mpORBvocabulary;
    ORBextractor *mpORBextractorLeft, *mpORBextractorRight;

    double mTimeStamp;
    cv::Mat mK;
    static float fx, fy, cx, cy, invfx, invfy;
    cv::Mat mDistCoef;
    float mbf; // stereo baseline multiplied by fx
    float mb;  // stereo baseline in meters
    float mThDepth;

    int N; // number of keypoints

    std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
    std::vector<cv::KeyPoint> mvKeysUn;
    std::vector<float> mvuRight; // negative value for monocular points
    std::vector<float> mvDepth;
    cv::Mat mDescriptors, mDescriptorsRight;

    int mnScaleLevels;
    float mfScaleFactor;
    float mfLogScaleFactor;
    std::vector<float> mvScaleFactors;
    std::vector<float> mvInvScaleFactors;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

    static bool mbInitialComputations;

private:
    // identity when the images are rectified (k1 == 0, all shipped YAMLs); reference Frame.cc:2049-2079
    void UndistortKeyPoints();
};

} // namespace ORB_SLAM2

#endif
