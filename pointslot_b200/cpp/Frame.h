// Frame.h — the stereo-construction path of PointSLOT's Frame (reference include/Frame.h:65, 94, 134,
// 232-273, 303-325; src/Frame.cc:144-200, 1636-1664, 1808-1861, 2027-2038, 2049-2079, 2112-2140, 2142-2316) as a
// drop-in: same constructor signature, same member names, same outputs (mvKeys, mvKeysRight, mvKeysUn,
// mDescriptors, mDescriptorsRight, mvuRight, mvDepth, mGrid, N, mbf, mb).  Everything else in the reference's Frame
// (object modes, BoW, pose) consumes these members and is out of scope; in a PointSLOT tree only the member
// functions below are replaced.
#ifndef FRAME_H
#define FRAME_H

#include <cstddef>
#include <vector>

#include "ORBextractor.h"
#include "compat/opencv_min.h"

namespace ORB_SLAM2
{
#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64

class ORBVocabulary; // opaque here: the stereo path never touches it
class MapPoint;      // only pointers are stored

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

    // Compute the cell of a keypoint (return false if outside the grid) (reference Frame.cc:2027-2038)
    bool PosInGrid(const cv::KeyPoint &kp, int &posX, int &posY);

    // Features of mGrid inside the square window |dx| < r, |dy| < r (reference Frame.cc:1808-1861)
    std::vector<size_t> GetFeaturesInArea(const float &x, const float &y, const float &r, const int minLevel = -1,
                                          const int maxLevel = -1) const;
    // Extension: many windows at once on the device, against the grid AssignFeaturesToGrid left there; same
    // results as GetFeaturesInArea per window (the search the projection matchers issue per map point).
    struct AreaQuery { float x, y, r; int minLevel, maxLevel; };
    std::vector<std::vector<size_t>> GetFeaturesInAreas(const std::vector<AreaQuery> &queries) const;

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

    // MapPoints associated to keypoints, NULL pointer if no association.
    std::vector<MapPoint *> mvpMapPoints;

    int mnScaleLevels;
    float mfScaleFactor;
    float mfLogScaleFactor;
    std::vector<float> mvScaleFactors;
    std::vector<float> mvInvScaleFactors;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

    // Keypoints are assigned to cells in a grid to reduce matching complexity when projecting MapPoints.
    static float mfGridElementWidthInv;
    static float mfGridElementHeightInv;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];

    // Undistorted Image Bounds (computed once).
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;

    static bool mbInitialComputations;

private:
    // identity when the images are rectified (k1 == 0, all shipped YAMLs); reference Frame.cc:2049-2079
    void UndistortKeyPoints();
    // Computes image bounds for the undistorted image (reference Frame.cc:2112-2140)
    void ComputeImageBounds(const cv::Mat &imLeft);
    // Assign keypoints to the grid for speed up feature matching (reference Frame.cc:1636-1655); on the device
    void AssignFeaturesToGrid();
};

} // namespace ORB_SLAM2

#endif
