// ChainBundle.h — host mirror of the reference bundle adjuster (include/mcptam/ChainBundle.h:99-224) with the
// same public signature, implemented on the B200 path through the C ABI (include/mcptam_b200.h).
// AddPose/AddPoint/AddMeas only collect flat arrays; the first Compute() uploads them (mcp_ba_load) and
// subsequent Compute() calls on the same object continue from the current estimate, as the reference's
// two-step adjustment does (src/BundleAdjusterMulti.cc:210-223).  Device buffers are pooled across objects
// (the reference constructs one ChainBundle per BundleAdjust call on the stack, :75).
#pragma once

#include <string>
#include <tuple>
#include <vector>

#include "TaylorCamera.h"

namespace mcp_host {

class ChainBundle {
 public:
  ChainBundle(TaylorCameraMap& cameraModels, bool bUseRobust, bool bUseTukey, bool bVerbose);
  ~ChainBundle();
  int AddPose(SE3 se3PoseFromRef, bool bFixed);                                        // src/ChainBundle.cc:1198
  int AddPoint(Vector<3> v3PointInCam, std::vector<int> vCams, bool bFixed);           // :1211
  void AddMeas(std::vector<int> vCams, int nPointIdx, Vector<2> v2Pos, double dNoiseSigmaSquared, std::string cameraName);  // :1239
  // The same measurement without the per-call vector / string copies of the reference signature: the two pose ids of the
  // observing chain and the index CameraIndex(name) returned for the camera.  BundleAdjusterCuda marshals through these.
  int CameraIndex(const std::string& cameraName);
  int AddPoint(const Vector<3>& v3Pos, int nPose0, int nPose1, bool bFixed);      // chain {nPose0[, nPose1]} (nPose1 = -1: one link), no vector argument
  void AddMeas(int nBasePoseId, int nCamPoseId, int nPointIdx, const Vector<2>& v2Pos, double dNoiseSigmaSquared, int nCameraIndex);
  // n measurements of one keyframe (same chain, same camera) in one call: what BundleAdjusterCuda::Marshal's parallel pass hands over
  void AddMeasBlock(int nBasePoseId, int nCamPoseId, int nCameraIndex, size_t n, const int* pPointIds, const double* pXy, const double* pNoise);
  // the same in two steps, for callers that fill disjoint blocks from several threads: GrowMeas(n) makes room for n more
  // measurements and returns the position of the first, FillMeasBlock writes one keyframe's block at a position inside it
  size_t GrowMeas(size_t n);
  void FillMeasBlock(size_t nAt, int nBasePoseId, int nCamPoseId, int nCameraIndex, size_t n, const int* pPointIds, const double* pXy, const double* pNoise);
  void Reserve(size_t nPoses, size_t nPoints, size_t nMeas);
  int Compute(bool* pAbortSignal, int nNumIter = snMaxIterations, double dUserLambda = -1);   // :1305
  bool Converged() { return mbConverged; }
  int TotalIterations() { return mnTotalIterations; }
  Vector<3> GetPoint(int n);
  SE3 GetPose(int n);
  std::vector<std::tuple<int, int, std::string> > GetOutlierMeasurements();
  double GetSigmaSquared() { return mdSigmaSquared; }
  double GetMeanChiSquared() { return mdMeanChiSquared; }
  double GetMaxCov() { return mdLastMaxCov; }
  double GetLambda() { return mdLambda; }
  double LastGpuMs() const { return mdGpuMs; }

  static int snMaxIterations;             // :1132
  static int snMaxTrialsAfterFailure;     // :1133
  static double sdUpdatePercentConvergenceLimit, sdUpdateRMSConvergenceLimit, sdMinMEstimatorSigma;   // :1134-1136

 protected:
  int Upload();
  void Fetch();
  TaylorCameraMap& mmCameraModels;
  bool mbUseRobust, mbUseTukey, mbVerbose;
  int mnCurrId = 1;                        // ids are one shared counter starting at 1 (:1145)
  std::vector<int> mvIdKind, mvIdIndex;    // id -> (0 pose / 1 point), index
  std::vector<int> mvPtId;                 // point index -> id
  McpBaConfig mConfig;                     // configuration the device handle was created with
  std::vector<double> mvPoseRt, mvPtXyz, mvMeasXy, mvMeasNoise;
  std::vector<uint8_t> mvPoseFixed, mvPtFixed;
  std::vector<int32_t> mvPtChain, mvMeasChain, mvMeasPt, mvMeasCam, mvMeasFirstId;
  std::vector<std::string> mvCamNames;     // camera index -> name
  std::vector<int> mvMeasCamName;
  McpBa* mpHandle = nullptr;
  bool mbHandleFailed = false;        // the last device call failed: the handle is destroyed instead of pooled
  bool mbUploaded = false, mbConverged = false;
  int mnTotalIterations = 0;
  double mdSigmaSquared = 0, mdMeanChiSquared = 0, mdLastMaxCov = 1.7976931348623157e308, mdLambda = 0, mdGpuMs = 0;
  std::vector<std::tuple<int, int, std::string> > mvOutlierMeasurementIdx;
};

}  // namespace mcp_host
