// Shape-only stand-in for srrg_hbst (tests/stubs/README.md).
#pragma once
#include <cstdint>
#include <map>
#include <vector>
#include <opencv2/opencv.hpp>
namespace srrg_hbst {
template <class ObjectT, unsigned Bits> class BinaryMatchable {
 public:
  typedef ObjectT ObjectType;
  BinaryMatchable(ObjectT, const cv::Mat&, uint64_t = 0);
  ObjectT object;
};
template <class MatchableT, class RealT> class BinaryNode {
 public:
  typedef MatchableT Matchable; typedef std::vector<MatchableT*> MatchableVector;
};
template <class NodeT> class BinaryTree {
 public:
  typedef NodeT Node; typedef typename NodeT::Matchable Matchable; typedef typename NodeT::MatchableVector MatchableVector;
  struct Match { const Matchable* matchable_query; const Matchable* matchable_reference; typename Matchable::ObjectType object_query, object_reference; double distance; };
  typedef std::vector<Match> MatchVector; typedef std::map<uint64_t, MatchVector> MatchVectorMap;
};
}  // namespace srrg_hbst
