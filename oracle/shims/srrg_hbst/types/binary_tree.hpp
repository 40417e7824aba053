// oracle/shims -- stand-in for srrg_hbst (relocalisation; outside the hot path).
// TEST INFRASTRUCTURE ONLY.  The hot path only creates, stores and deletes matchables (src/types/landmark.cpp:24-63,
// local_map.cpp:55-60); the tree search itself belongs to the relocaliser, which oracle/_ref does not build.
#pragma once
#include <cstdint>
#include <map>
#include <vector>
#include <opencv2/opencv.hpp>
namespace srrg_hbst {
template <class ObjectT, unsigned Bits>
class BinaryMatchable {
 public:
  typedef ObjectT ObjectType;
  BinaryMatchable(ObjectT object_, const cv::Mat& descriptor_, uint64_t identifier_tree_ = 0)
      : object(object_), descriptor(descriptor_), identifier_tree(identifier_tree_) {}
  void setObjects(ObjectT o) { object = o; }   // src/types/landmark.cpp:181
  ObjectT object;
  cv::Mat descriptor;
  uint64_t identifier_tree;
};
template <class MatchableT, class RealT>
class BinaryNode {
 public:
  typedef MatchableT Matchable;
  typedef std::vector<MatchableT*> MatchableVector;
};
template <class NodeT>
class BinaryTree {
 public:
  typedef NodeT Node;
  typedef typename NodeT::Matchable Matchable;
  typedef typename NodeT::MatchableVector MatchableVector;
  struct Match {
    const Matchable* matchable_query;
    const Matchable* matchable_reference;
    typename Matchable::ObjectType object_query, object_reference;
    double distance;
  };
  typedef std::vector<Match> MatchVector;
  typedef std::map<uint64_t, MatchVector> MatchVectorMap;
};
}  // namespace srrg_hbst
