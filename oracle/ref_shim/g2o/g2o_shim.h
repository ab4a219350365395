// Minimal stand-in for the part of g2o (2013-era API) that src/ChainBundle.cc is written against, so that the reference's
// own vertices, edges, robust kernel, actions and ChainBundle::Compute run unmodified (TEST INFRASTRUCTURE, oracle/_ref).
// [3P] The optimiser is a DENSE restatement of g2o's SparseOptimizer::optimize + OptimizationAlgorithmLevenberg::solve
// (same control flow, +lambda on every diagonal entry, computeLambdaInit = 1e-5 max H_jj, computeScale = sum x (lambda x + b),
// rho = (chi - tempChi) / (scale + 1e-3), lambda schedule 1/3 .. 2/3, ni doubling) with a dense Cholesky of the full
// (poses + points) system in place of BlockSolverX + CHOLMOD: small problems only.
#pragma once
#include <Eigen/Core>
#include <cmath>
#include <iostream>
#include <limits>
#include <map>
#include <set>
#include <vector>

namespace g2o {
using namespace Eigen;

class HyperGraph {
public:
  class Vertex {
  public:
    Vertex() : _id(-1) {}
    virtual ~Vertex() {}
    int id() const { return _id; }
    virtual void setId(int i) { _id = i; }
  protected:
    int _id;
  };
  class Edge {
  public:
    virtual ~Edge() {}
    std::vector<Vertex*>& vertices() { return _vertices; }
    const std::vector<Vertex*>& vertices() const { return _vertices; }
    virtual void resize(size_t n) { _vertices.resize(n, 0); }
  protected:
    std::vector<Vertex*> _vertices;
  };
  virtual ~HyperGraph() {}
};

class HyperGraphAction {
public:
  class Parameters {};
  virtual ~HyperGraphAction() {}
  virtual HyperGraphAction* operator()(const HyperGraph* graph, Parameters* parameters = 0) = 0;
};

class RobustKernel {
public:
  virtual ~RobustKernel() {}
  virtual void robustify(double e2, Eigen::Vector3d& rho) const = 0;
};

class OptimizableGraph : public HyperGraph {
public:
  class Vertex : public HyperGraph::Vertex {
  public:
    Vertex() : _fixed(false), _marginalized(false), _hessianIndex(-1) {}
    bool fixed() const { return _fixed; }
    void setFixed(bool f) { _fixed = f; }
    bool marginalized() const { return _marginalized; }
    void setMarginalized(bool m) { _marginalized = m; }
    int hessianIndex() const { return _hessianIndex; }
    void setHessianIndex(int i) { _hessianIndex = i; }
    virtual int dimension() const = 0;
    void oplus(const double* v) { oplusImpl(v); }
    void setToOrigin() { setToOriginImpl(); }
    virtual void push() = 0;
    virtual void pop() = 0;
    virtual void discardTop() = 0;
    virtual void oplusImpl(const double* v) = 0;
    virtual void setToOriginImpl() = 0;
    virtual bool read(std::istream& is) = 0;
    virtual bool write(std::ostream& os) const = 0;
  protected:
    bool _fixed, _marginalized;
    int _hessianIndex;
  };
  class Edge : public HyperGraph::Edge {
  public:
    Edge() : _robustKernel(0) {}
    virtual ~Edge() { delete _robustKernel; }
    RobustKernel* robustKernel() const { return _robustKernel; }
    void setRobustKernel(RobustKernel* k) { delete _robustKernel; _robustKernel = k; }
    virtual void computeError() = 0;
    virtual void linearizeOplus() = 0;
    virtual double chi2() const = 0;
    virtual int dimension() const = 0;
    virtual const double* errorData() const = 0;
    virtual const double* informationData() const = 0;
    virtual const double* jacobianData(int i) const = 0;       // D x dim(vertex i), row-major
    virtual bool read(std::istream& is) = 0;
    virtual bool write(std::ostream& os) const = 0;
  protected:
    RobustKernel* _robustKernel;
  };
  typedef std::vector<Vertex*> VertexContainer;
  typedef std::vector<Edge*> EdgeContainer;
  virtual ~OptimizableGraph()
  {
    for (size_t i = 0; i < _edges.size(); i++) delete _edges[i];
    for (std::map<int, Vertex*>::iterator it = _vertices.begin(); it != _vertices.end(); ++it) delete it->second;
  }
  bool addVertex(Vertex* v) { _vertices[v->id()] = v; return true; }
  bool addEdge(Edge* e) { _edges.push_back(e); return true; }
  Vertex* vertex(int id) { std::map<int, Vertex*>::iterator it = _vertices.find(id); return it == _vertices.end() ? 0 : it->second; }
  const Vertex* vertex(int id) const { std::map<int, Vertex*>::const_iterator it = _vertices.find(id); return it == _vertices.end() ? 0 : it->second; }
protected:
  std::map<int, Vertex*> _vertices;
  std::vector<Edge*> _edges;
};

template <int D, class T> class BaseVertex : public OptimizableGraph::Vertex {
public:
  static const int Dimension = D;
  typedef T EstimateType;
  const T& estimate() const { return _estimate; }
  void setEstimate(const T& e) { _estimate = e; }
  virtual int dimension() const { return D; }
  virtual void push() { _backup.push_back(_estimate); }
  virtual void pop() { _estimate = _backup.back(); _backup.pop_back(); }
  virtual void discardTop() { _backup.pop_back(); }
protected:
  T _estimate;
  std::vector<T> _backup;
};

template <int D, class E> class BaseMultiEdge : public OptimizableGraph::Edge {
public:
  static const int Dimension = D;
  typedef E Measurement;
  typedef Eigen::Matrix<double, D, Eigen::Dynamic> JacobianType;
  typedef Eigen::Matrix<double, D, 1> ErrorVector;
  typedef Eigen::Matrix<double, D, D> InformationType;
  BaseMultiEdge() : _dimension(D) {}
  virtual void resize(size_t n) { OptimizableGraph::Edge::resize(n); _jacobianOplus.resize(n); }
  const E& measurement() const { return _measurement; }
  void setMeasurement(const E& m) { _measurement = m; }
  const InformationType& information() const { return _information; }
  InformationType& information() { return _information; }
  const ErrorVector& error() const { return _error; }
  virtual double chi2() const { return _error.dot(_information * _error); }
  virtual int dimension() const { return D; }
  virtual const double* errorData() const { return _error.data(); }
  virtual const double* informationData() const { return _information.data(); }
  virtual const double* jacobianData(int i) const { return _jacobianOplus[i].data(); }
protected:
  int _dimension;
  E _measurement;
  InformationType _information;
  ErrorVector _error;
  std::vector<JacobianType, Eigen::aligned_allocator<JacobianType> > _jacobianOplus;
};

class Solver {
public:
  Solver() {}
  virtual ~Solver() {}
  const double* x() const { return _x.empty() ? 0 : &_x[0]; }
  int vectorSize() const { return (int)_x.size(); }
  std::vector<double> _x, _b, _H;     // dense system of the last buildSystem (row-major, dimension vectorSize())
};
template <class M> class LinearSolverCholmod { public: LinearSolverCholmod() {} };
class BlockSolverX : public Solver {
public:
  typedef Eigen::MatrixXd PoseMatrixType;
  typedef LinearSolverCholmod<PoseMatrixType> LinearSolverType;
  BlockSolverX(LinearSolverType* ls) : _ls(ls) {}
  ~BlockSolverX() { delete _ls; }
private:
  LinearSolverType* _ls;
};
template <class M> class SparseBlockMatrix {
public:
  std::vector<M> blocks;
  const M* block(int r, int c) const { (void)c; return &blocks[r]; }
};

class SparseOptimizer;
class OptimizationAlgorithm {
public:
  enum SolverResult { Terminate = 2, OK = 1, Fail = -1 };
  virtual ~OptimizationAlgorithm() {}
  virtual SolverResult solve(int iteration, bool online = false) = 0;
  void setOptimizer(SparseOptimizer* o) { _optimizer = o; }
protected:
  SparseOptimizer* _optimizer;
};
class OptimizationAlgorithmWithHessian : public OptimizationAlgorithm {
public:
  OptimizationAlgorithmWithHessian(Solver* s) : _solver(s) {}
  ~OptimizationAlgorithmWithHessian() { delete _solver; }
  Solver* solver() { return _solver; }
protected:
  Solver* _solver;
};
class OptimizationAlgorithmLevenberg : public OptimizationAlgorithmWithHessian {
public:
  OptimizationAlgorithmLevenberg(Solver* s) : OptimizationAlgorithmWithHessian(s), _currentLambda(-1), _tau(1e-5), _goodStepLowerScale(1. / 3.), _goodStepUpperScale(2. / 3.), _ni(2), _levenbergIterations(0), _maxTrials(10), _userLambdaInit(0) {}
  void setMaxTrialsAfterFailure(int n) { _maxTrials = n; }
  void setUserLambdaInit(double l) { _userLambdaInit = l; }
  int levenbergIteration() { return _levenbergIterations; }
  double currentLambda() const { return _currentLambda; }
  virtual SolverResult solve(int iteration, bool online = false);
protected:
  double _currentLambda, _tau, _goodStepLowerScale, _goodStepUpperScale, _ni;
  int _levenbergIterations, _maxTrials;
  double _userLambdaInit;
};

class SparseOptimizer : public OptimizableGraph {
public:
  SparseOptimizer() : _algorithm(0), _forceStop(0), _verbose(false) {}
  ~SparseOptimizer() { delete _algorithm; }
  void setAlgorithm(OptimizationAlgorithm* a) { _algorithm = a; a->setOptimizer(this); }
  OptimizationAlgorithm* solver() { return _algorithm; }
  void setVerbose(bool v) { _verbose = v; }
  void setForceStopFlag(bool* f) { _forceStop = f; }
  bool terminate() const { return _forceStop ? *_forceStop : false; }
  bool addPreIterationAction(HyperGraphAction* a) { _pre.push_back(a); return true; }
  bool addPostIterationAction(HyperGraphAction* a) { _post.push_back(a); return true; }
  bool addComputeErrorAction(HyperGraphAction* a) { _err.push_back(a); return true; }
  const EdgeContainer& activeEdges() const { return _activeEdges; }
  const VertexContainer& activeVertices() const { return _activeVertices; }
  bool initializeOptimization(int level = 0)
  {
    (void)level;
    _activeVertices.clear(); _activeEdges = _edges; _index.clear();
    int off = 0;
    for (std::map<int, Vertex*>::iterator it = _vertices.begin(); it != _vertices.end(); ++it) {
      _activeVertices.push_back(it->second);
      if (!it->second->fixed()) { it->second->setHessianIndex((int)_index.size()); _index.push_back(it->second); _offset.resize(_index.size()); _offset[_index.size() - 1] = off; off += it->second->dimension(); }
      else it->second->setHessianIndex(-1);
    }
    _dim = off;
    return true;
  }
  void computeActiveErrors()
  {
    for (size_t i = 0; i < _err.size(); i++) (*_err[i])(this);
    for (size_t i = 0; i < _activeEdges.size(); i++) _activeEdges[i]->computeError();
  }
  double activeRobustChi2() const
  {
    double chi = 0;
    Eigen::Vector3d rho;
    for (size_t i = 0; i < _activeEdges.size(); i++) {
      const Edge* e = _activeEdges[i];
      if (e->robustKernel()) { e->robustKernel()->robustify(e->chi2(), rho); chi += rho[0]; }
      else chi += e->chi2();
    }
    return chi;
  }
  void push() { for (size_t i = 0; i < _activeVertices.size(); i++) _activeVertices[i]->push(); }
  void pop() { for (size_t i = 0; i < _activeVertices.size(); i++) _activeVertices[i]->pop(); }
  void discardTop() { for (size_t i = 0; i < _activeVertices.size(); i++) _activeVertices[i]->discardTop(); }
  void update(const double* x) { for (size_t i = 0; i < _index.size(); i++) _index[i]->oplus(x + _offset[i]); }
  int optimize(int iterations, bool online = false)
  {
    (void)online;
    int cj = 0;
    bool ok = true;
    for (int i = 0; i < iterations && !terminate() && ok; i++) {
      for (size_t k = 0; k < _pre.size(); k++) (*_pre[k])(this);
      const OptimizationAlgorithm::SolverResult r = _algorithm->solve(i, false);
      ok = (r == OptimizationAlgorithm::OK);
      ++cj;
      for (size_t k = 0; k < _post.size(); k++) (*_post[k])(this);
    }
    if (!ok) return 0;
    return cj;
  }
  // dense normal equations of the active graph: g2o BaseMultiEdge::constructQuadraticForm / computeQuadraticForm
  void buildSystem(Solver* s)
  {
    const int n = _dim;
    s->_H.assign((size_t)n * n, 0.0); s->_b.assign((size_t)n, 0.0); s->_x.assign((size_t)n, 0.0);
    Eigen::Vector3d rho;
    for (size_t q = 0; q < _activeEdges.size(); q++) {
      Edge* e = _activeEdges[q];
      e->linearizeOplus();
      const int D = e->dimension();
      const double* err = e->errorData(); const double* info = e->informationData();
      double w = 1.0;
      if (e->robustKernel()) { e->robustKernel()->robustify(e->chi2(), rho); w = rho[1]; }
      std::vector<double> omega((size_t)D * D), omega_r((size_t)D, 0.0);
      for (int a = 0; a < D; a++) { for (int b = 0; b < D; b++) { omega[a * D + b] = w * info[a * D + b]; omega_r[a] -= info[a * D + b] * err[b]; } omega_r[a] *= w; }
      const std::vector<HyperGraph::Vertex*>& vs = e->vertices();
      for (size_t i = 0; i < vs.size(); i++) {
        Vertex* vi = static_cast<Vertex*>(vs[i]);
        if (vi->fixed()) continue;
        const int di = vi->dimension(), oi = _offset[vi->hessianIndex()];
        const double* Ji = e->jacobianData((int)i);
        std::vector<double> A((size_t)di * D);                      // J_i^T omega
        for (int a = 0; a < di; a++) for (int b = 0; b < D; b++) { double t = 0; for (int c = 0; c < D; c++) t += Ji[c * di + a] * omega[c * D + b]; A[a * D + b] = t; }
        for (int a = 0; a < di; a++) { double t = 0; for (int b = 0; b < D; b++) t += Ji[b * di + a] * omega_r[b]; s->_b[oi + a] += t; }
        for (size_t j = i; j < vs.size(); j++) {
          Vertex* vj = static_cast<Vertex*>(vs[j]);
          if (vj->fixed()) continue;
          const int dj = vj->dimension(), oj = _offset[vj->hessianIndex()];
          const double* Jj = e->jacobianData((int)j);
          for (int a = 0; a < di; a++) for (int b = 0; b < dj; b++) {
            double t = 0;
            for (int c = 0; c < D; c++) t += A[a * D + c] * Jj[c * dj + b];
            s->_H[(size_t)(oi + a) * n + oj + b] += t;
            if (oi != oj) s->_H[(size_t)(oj + b) * n + oi + a] += t;
          }
        }
      }
    }
  }
  // median-ready marginals: diagonal blocks of H^-1 of the last (undamped) system for the listed vertices
  bool computeMarginals(SparseBlockMatrix<Eigen::MatrixXd>& spinv, const VertexContainer& vertices)
  {
    Solver* s = static_cast<OptimizationAlgorithmWithHessian*>(_algorithm)->solver();
    const int n = _dim;
    if ((int)s->_H.size() != n * n || n == 0) return false;
    std::vector<double> L(s->_H);
    for (int j = 0; j < n; j++) {
      double dj = L[(size_t)j * n + j];
      for (int k = 0; k < j; k++) dj -= L[(size_t)j * n + k] * L[(size_t)j * n + k];
      if (!(dj > 0)) return false;
      dj = std::sqrt(dj); L[(size_t)j * n + j] = dj;
      for (int i = j + 1; i < n; i++) { double t = L[(size_t)i * n + j]; for (int k = 0; k < j; k++) t -= L[(size_t)i * n + k] * L[(size_t)j * n + k]; L[(size_t)i * n + j] = t / dj; }
    }
    spinv.blocks.clear();
    for (size_t v = 0; v < vertices.size(); v++) {
      const int d = vertices[v]->dimension(), o = _offset[vertices[v]->hessianIndex()];
      Eigen::MatrixXd B(d, d);
      for (int c = 0; c < d; c++) {
        std::vector<double> y((size_t)n, 0.0), x((size_t)n, 0.0);
        for (int i = 0; i < n; i++) { double t = (i == o + c) ? 1.0 : 0.0; for (int k = 0; k < i; k++) t -= L[(size_t)i * n + k] * y[k]; y[i] = t / L[(size_t)i * n + i]; }
        for (int i = n - 1; i >= 0; i--) { double t = y[i]; for (int k = i + 1; k < n; k++) t -= L[(size_t)k * n + i] * x[k]; x[i] = t / L[(size_t)i * n + i]; }
        for (int r = 0; r < d; r++) B(r, c) = x[o + r];
      }
      spinv.blocks.push_back(B);
    }
    return true;
  }
  int dim() const { return _dim; }
private:
  OptimizationAlgorithm* _algorithm;
  bool* _forceStop;
  bool _verbose;
  std::vector<HyperGraphAction*> _pre, _post, _err;
  EdgeContainer _activeEdges;
  VertexContainer _activeVertices, _index;
  std::vector<int> _offset;
  int _dim;
};

inline OptimizationAlgorithm::SolverResult OptimizationAlgorithmLevenberg::solve(int iteration, bool online)
{
  (void)online;
  Solver* s = _solver;
  _optimizer->computeActiveErrors();
  double currentChi = _optimizer->activeRobustChi2();
  double tempChi = currentChi;
  _optimizer->buildSystem(s);
  const int n = _optimizer->dim();
  if (iteration == 0) {
    if (_userLambdaInit > 0) _currentLambda = _userLambdaInit;
    else { double mx = 0; for (int j = 0; j < n; j++) mx = std::max(std::fabs(s->_H[(size_t)j * n + j]), mx); _currentLambda = _tau * mx; }
    _ni = 2;
  }
  double rho = 0;
  int& qmax = _levenbergIterations;
  qmax = 0;
  do {
    _optimizer->push();
    // (H + lambda I) x = b by dense Cholesky
    std::vector<double> L(s->_H);
    for (int j = 0; j < n; j++) L[(size_t)j * n + j] += _currentLambda;
    bool ok2 = true;
    for (int j = 0; j < n && ok2; j++) {
      double dj = L[(size_t)j * n + j];
      for (int k = 0; k < j; k++) dj -= L[(size_t)j * n + k] * L[(size_t)j * n + k];
      if (!(dj > 0) || !std::isfinite(dj)) { ok2 = false; break; }
      dj = std::sqrt(dj); L[(size_t)j * n + j] = dj;
      for (int i = j + 1; i < n; i++) { double t = L[(size_t)i * n + j]; for (int k = 0; k < j; k++) t -= L[(size_t)i * n + k] * L[(size_t)j * n + k]; L[(size_t)i * n + j] = t / dj; }
    }
    if (ok2) {
      std::vector<double> y((size_t)n);
      for (int i = 0; i < n; i++) { double t = s->_b[i]; for (int k = 0; k < i; k++) t -= L[(size_t)i * n + k] * y[k]; y[i] = t / L[(size_t)i * n + i]; }
      for (int i = n - 1; i >= 0; i--) { double t = y[i]; for (int k = i + 1; k < n; k++) t -= L[(size_t)k * n + i] * s->_x[k]; s->_x[i] = t / L[(size_t)i * n + i]; }
    } else s->_x.assign((size_t)n, 0.0);
    _optimizer->update(s->x());
    _optimizer->computeActiveErrors();
    tempChi = _optimizer->activeRobustChi2();
    if (!ok2) tempChi = std::numeric_limits<double>::max();
    rho = (currentChi - tempChi);
    double scale = 0;
    for (int j = 0; j < n; j++) scale += s->_x[j] * (_currentLambda * s->_x[j] + s->_b[j]);
    scale += 1e-3;
    rho /= scale;
    if (rho > 0 && std::isfinite(tempChi)) {
      double alpha = 1. - std::pow((2 * rho - 1), 3);
      alpha = (std::min)(alpha, _goodStepUpperScale);
      const double scaleFactor = (std::max)(_goodStepLowerScale, alpha);
      _currentLambda *= scaleFactor;
      _ni = 2;
      currentChi = tempChi;
      _optimizer->discardTop();
    } else {
      _currentLambda *= _ni;
      _ni *= 2;
      _optimizer->pop();
    }
    qmax++;
  } while (rho < 0 && qmax < _maxTrials && !_optimizer->terminate());
  if (qmax == _maxTrials || rho == 0) return Terminate;
  return OK;
}

}  // namespace g2o
