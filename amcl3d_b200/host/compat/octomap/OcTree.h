// compat stand-in for <octomap/OcTree.h>: a small, self-contained occupancy octree that speaks
// octomap's published stream formats (".bt" binary maximum-likelihood trees and ".ot" full
// log-odds trees) and exposes the members amcl3d touches
// (reference use sites: amcl3d/src/PointCloudTools.cpp:35-42 and :56-75).
//
// Written from the published format/semantics, not from octomap sources:
//  * depth-16 tree, key origin 32768, leaf centre = (key - 32768 + 0.5) * resolution,
//    a node at depth d spans 2^(16-d) finest voxels and is centred on its key block;
//  * ".bt": text header ("# Octomap OcTree binary file", id/size/res/data) followed by a
//    pre-order stream of 2 bytes per inner node, 2 bits per child
//    (00 unknown, 01 occupied leaf, 10 free leaf, 11 inner node) with bit 0 = LSB, child i in
//    bits (2i, 2i+1) of byte i/4;
//  * ".ot": text header ("# Octomap OcTree file") followed by pre-order
//    (float log-odds, child-mask byte) records;
//  * occupancy threshold log-odds 0, clamping log-odds [-2, 3.5] (octomap defaults);
//  * metric bounds are taken over ALL leaves (free ones included), per leaf
//    centre -/+ half its size, in double.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <istream>
#include <limits>
#include <ostream>
#include <sstream>
#include <string>
#include <vector>

namespace octomap
{
class OcTreeNode
{
public:
  float getLogOdds() const { return value_; }
  void setLogOdds(float v) { value_ = v; }
  double getOccupancy() const { return 1.0 - 1.0 / (1.0 + std::exp(static_cast<double>(value_))); }

private:
  friend class OcTree;
  float value_{ 0.f };
  uint32_t kids_{ 0 };  // index of this node's 8-slot child block (0 = leaf)
  uint8_t mask_{ 0 };   // bit i set <=> child i exists
};

class AbstractOcTree
{
public:
  virtual ~AbstractOcTree() {}
  virtual std::size_t size() const = 0;
  virtual double getResolution() const = 0;
  virtual std::string getTreeType() const = 0;
  // Reads a ".ot" stream; returns a heap-allocated tree or nullptr.  Defined after OcTree.
  static inline AbstractOcTree* read(const std::string& filename);
};

class OcTree : public AbstractOcTree
{
public:
  static const unsigned kDepth = 16;
  static const int kKeyOrigin = 32768;

  explicit OcTree(double resolution) : resolution_(resolution) { clear(); }
  ~OcTree() override {}

  std::string getTreeType() const override { return "OcTree"; }
  double getResolution() const override { return resolution_; }
  void setResolution(double r) { resolution_ = r; }
  std::size_t size() const override { return tree_size_; }
  unsigned getTreeDepth() const { return kDepth; }
  double getNodeSize(unsigned depth) const { return resolution_ * static_cast<double>(1u << (kDepth - depth)); }

  void clear()
  {
    nodes_.clear();
    nodes_.resize(8);  // block 0 is a sentinel so that kids_ == 0 can mean "no children"
    has_root_ = false;
    tree_size_ = 0;
    bounds_valid_ = false;
  }

  bool isNodeOccupied(const OcTreeNode& n) const { return n.getLogOdds() >= occ_thres_log_; }
  bool isNodeOccupied(const OcTreeNode* n) const { return n->getLogOdds() >= occ_thres_log_; }

  // ---------------------------------------------------------------- building (used by the synthetic map writers)
  // Marks the cell of size 2^(16-depth) voxels containing finest-voxel key (kx,ky,kz) as an occupied / free leaf.
  void setLeaf(uint32_t kx, uint32_t ky, uint32_t kz, unsigned depth, bool occupied)
  {
    if (!has_root_)
    {
      root_ = OcTreeNode();
      has_root_ = true;
    }
    OcTreeNode* n = &root_;
    for (unsigned d = 0; d < depth; ++d)
    {
      const unsigned bit = kDepth - 1 - d;
      const unsigned pos = ((kx >> bit) & 1u) | (((ky >> bit) & 1u) << 1) | (((kz >> bit) & 1u) << 2);
      n = &childOrCreate(n, pos);
    }
    n->kids_ = 0;
    n->mask_ = 0;
    n->value_ = occupied ? clamp_max_ : clamp_min_;
    bounds_valid_ = false;
    size_valid_ = false;
  }
  // Metric-coordinate convenience: finest-level leaf containing (x, y, z).
  void updateNode(double x, double y, double z, bool occupied)
  {
    setLeaf(coordToKey(x), coordToKey(y), coordToKey(z), kDepth, occupied);
  }
  uint32_t coordToKey(double c) const { return static_cast<uint32_t>(static_cast<int>(std::floor(c / resolution_)) + kKeyOrigin); }
  // Recomputes inner-node values (max over children) and the node count after edits.
  void updateInnerOccupancy()
  {
    if (has_root_)
      updateInnerRecurs(root_);
    tree_size_ = has_root_ ? countRecurs(root_) : 0;
    size_valid_ = true;
  }

  // ---------------------------------------------------------------- metric bounds
  void getMetricMin(double& x, double& y, double& z)
  {
    ensureBounds();
    x = min_[0];
    y = min_[1];
    z = min_[2];
  }
  void getMetricMax(double& x, double& y, double& z)
  {
    ensureBounds();
    x = max_[0];
    y = max_[1];
    z = max_[2];
  }

  // ---------------------------------------------------------------- leaf iteration (pre-order, child 0 first)
  class leaf_iterator
  {
  public:
    leaf_iterator() : tree_(nullptr) {}
    leaf_iterator(const OcTree* tree, bool at_begin) : tree_(tree)
    {
      if (at_begin && tree_->has_root_)
      {
        Frame f;
        f.node = &tree_->root_;
        f.key[0] = f.key[1] = f.key[2] = kKeyOrigin;
        f.depth = 0;
        stack_.push_back(f);
        descend();
      }
    }
    leaf_iterator& operator++()
    {
      if (!stack_.empty())
      {
        stack_.pop_back();
        descend();
      }
      return *this;
    }
    leaf_iterator operator++(int)
    {
      leaf_iterator r = *this;
      ++(*this);
      return r;
    }
    bool operator==(const leaf_iterator& o) const
    {
      if (stack_.empty() || o.stack_.empty())
        return stack_.empty() == o.stack_.empty();
      return stack_.back().node == o.stack_.back().node;
    }
    bool operator!=(const leaf_iterator& o) const { return !(*this == o); }
    bool operator!=(std::nullptr_t) const { return !stack_.empty(); }
    bool operator==(std::nullptr_t) const { return stack_.empty(); }
    const OcTreeNode& operator*() const { return *stack_.back().node; }
    const OcTreeNode* operator->() const { return stack_.back().node; }
    unsigned getDepth() const { return stack_.back().depth; }
    double getSize() const { return tree_->getNodeSize(stack_.back().depth); }
    double getX() const { return tree_->keyToCoord(stack_.back().key[0], stack_.back().depth); }
    double getY() const { return tree_->keyToCoord(stack_.back().key[1], stack_.back().depth); }
    double getZ() const { return tree_->keyToCoord(stack_.back().key[2], stack_.back().depth); }

  private:
    struct Frame
    {
      const OcTreeNode* node;
      uint32_t key[3];
      unsigned depth;
    };
    // Expand the top of the stack until it is a leaf (a node without children).
    void descend()
    {
      while (!stack_.empty() && stack_.back().node->kids_ != 0)
      {
        const Frame top = stack_.back();
        stack_.pop_back();
        const unsigned cd = top.depth + 1;
        const uint32_t off = static_cast<uint32_t>(kKeyOrigin) >> cd;
        for (int i = 7; i >= 0; --i)
        {
          if (!((top.node->mask_ >> i) & 1u))
            continue;
          Frame f;
          f.node = &tree_->nodes_[static_cast<std::size_t>(top.node->kids_) * 8 + i];
          f.depth = cd;
          for (int a = 0; a < 3; ++a)
            f.key[a] = ((i >> a) & 1) ? top.key[a] + off : top.key[a] - off - (off ? 0u : 1u);
          stack_.push_back(f);
        }
      }
    }
    const OcTree* tree_;
    std::vector<Frame> stack_;
  };
  leaf_iterator begin_leafs() const { return leaf_iterator(this, true); }
  leaf_iterator end_leafs() const { return leaf_iterator(this, false); }

  double keyToCoord(uint32_t key, unsigned depth) const
  {
    if (depth == 0)
      return 0.0;
    if (depth == kDepth)
      return (static_cast<double>(static_cast<int>(key) - kKeyOrigin) + 0.5) * resolution_;
    const double cells = static_cast<double>(1u << (kDepth - depth));
    return (std::floor((static_cast<double>(key) - static_cast<double>(kKeyOrigin)) / cells) + 0.5) * getNodeSize(depth);
  }

  // ---------------------------------------------------------------- ".bt" streams
  bool readBinary(const std::string& filename)
  {
    std::ifstream f(filename.c_str(), std::ios::in | std::ios::binary);
    if (!f.is_open())
      return false;
    return readBinary(f);
  }
  bool readBinary(std::istream& s)
  {
    std::string line;
    if (!std::getline(s, line) || line.compare(0, 28, "# Octomap OcTree binary file") != 0)
      return false;
    std::string id;
    uint64_t n = 0;
    double res = 0;
    if (!readHeaderFields(s, id, n, res) || id != "OcTree")
      return false;
    clear();
    resolution_ = res;
    if (n > 0)
    {
      root_ = OcTreeNode();
      has_root_ = true;
      if (!readBinaryRecurs(s, root_, 0))
      {
        clear();
        return false;
      }
      updateInnerOccupancy();
    }
    return tree_size_ == n;
  }
  bool writeBinary(const std::string& filename)
  {
    std::ofstream f(filename.c_str(), std::ios::out | std::ios::binary);
    if (!f.is_open())
      return false;
    if (!size_valid_)
      updateInnerOccupancy();
    f << "# Octomap OcTree binary file\n# (feel free to add / change comments, but leave the first line as it is!)\n#\n";
    f << "id OcTree\nsize " << tree_size_ << "\nres " << formatDouble(resolution_) << "\ndata\n";
    if (has_root_)
      writeBinaryRecurs(f, root_);
    return f.good();
  }

  // ---------------------------------------------------------------- ".ot" streams
  bool write(const std::string& filename)
  {
    std::ofstream f(filename.c_str(), std::ios::out | std::ios::binary);
    if (!f.is_open())
      return false;
    if (!size_valid_)
      updateInnerOccupancy();
    f << "# Octomap OcTree file\n# (feel free to add / change comments, but leave the first line as it is!)\n#\n";
    f << "id OcTree\nsize " << tree_size_ << "\nres " << formatDouble(resolution_) << "\ndata\n";
    if (has_root_)
      writeFullRecurs(f, root_);
    return f.good();
  }
  bool readData(std::istream& s)
  {
    root_ = OcTreeNode();
    has_root_ = true;
    if (!readFullRecurs(s, root_, 0))
    {
      clear();
      return false;
    }
    tree_size_ = countRecurs(root_);
    size_valid_ = true;
    bounds_valid_ = false;
    return true;
  }

  static bool readHeaderFields(std::istream& s, std::string& id, uint64_t& size, double& res)
  {
    id.clear();
    size = 0;
    res = 0.0;
    std::string token;
    bool done = false;
    while (s.good() && !done)
    {
      if (!(s >> token))
        break;
      if (token == "data")
      {
        done = true;
        skipLine(s);
      }
      else if (token[0] == '#')
        skipLine(s);
      else if (token == "id")
        s >> id;
      else if (token == "res")
        s >> res;
      else if (token == "size")
        s >> size;
      else
        skipLine(s);
    }
    if (!done || id.empty() || !(res > 0.0))
      return false;
    if (id == "1")
      id = "OcTree";  // legacy spelling
    return true;
  }

private:
  static void skipLine(std::istream& s)
  {
    char c;
    do
    {
      c = static_cast<char>(s.get());
    } while (s.good() && c != '\n');
  }
  static std::string formatDouble(double v)
  {
    std::ostringstream o;
    o.precision(17);
    o << v;
    // prefer the shortest representation that round-trips
    for (int p = 1; p < 17; ++p)
    {
      std::ostringstream t;
      t.precision(p);
      t << v;
      if (std::stod(t.str()) == v)
        return t.str();
    }
    return o.str();
  }
  OcTreeNode& childOrCreate(OcTreeNode*& n, unsigned pos)
  {
    if (n->kids_ == 0)
    {
      const std::size_t block = nodes_.size() / 8;
      const bool is_root = (n == &root_);
      const std::size_t self = is_root ? 0 : static_cast<std::size_t>(n - nodes_.data());
      nodes_.resize(nodes_.size() + 8);  // may reallocate: re-derive n
      n = is_root ? &root_ : &nodes_[self];
      n->kids_ = static_cast<uint32_t>(block);
      n->mask_ = 0;
    }
    n->mask_ = static_cast<uint8_t>(n->mask_ | (1u << pos));
    return nodes_[static_cast<std::size_t>(n->kids_) * 8 + pos];
  }
  OcTreeNode& child(const OcTreeNode& n, unsigned pos) { return nodes_[static_cast<std::size_t>(n.kids_) * 8 + pos]; }
  const OcTreeNode& child(const OcTreeNode& n, unsigned pos) const
  {
    return nodes_[static_cast<std::size_t>(n.kids_) * 8 + pos];
  }

  bool readBinaryRecurs(std::istream& s, OcTreeNode& node_in, unsigned depth)
  {
    if (depth >= kDepth)
      return false;
    unsigned char b[2];
    s.read(reinterpret_cast<char*>(b), 2);
    if (!s.good() && !s.eof())
      return false;
    if (s.gcount() != 2)
      return false;
    const unsigned bits = static_cast<unsigned>(b[0]) | (static_cast<unsigned>(b[1]) << 8);
    OcTreeNode* node = &node_in;
    const bool node_is_root = (node == &root_);
    const std::size_t self = node_is_root ? 0 : static_cast<std::size_t>(node - nodes_.data());
    node->value_ = clamp_max_;
    unsigned inner = 0;
    for (unsigned i = 0; i < 8; ++i)
    {
      const unsigned lo = (bits >> (2 * i)) & 1u, hi = (bits >> (2 * i + 1)) & 1u;
      if (!lo && !hi)
        continue;
      OcTreeNode& c = childOrCreate(node, i);
      c = OcTreeNode();
      if (lo && !hi)
        c.value_ = clamp_min_;  // free leaf
      else if (!lo && hi)
        c.value_ = clamp_max_;  // occupied leaf
      else
        inner |= 1u << i;
    }
    for (unsigned i = 0; i < 8; ++i)
    {
      if (!((inner >> i) & 1u))
        continue;
      // nodes_ may have been reallocated by deeper recursion: always re-derive pointers from indices
      OcTreeNode& me = node_is_root ? root_ : nodes_[self];
      const std::size_t ci = static_cast<std::size_t>(me.kids_) * 8 + i;
      if (!readBinaryRecurs(s, nodes_[ci], depth + 1))
        return false;
    }
    return true;
  }
  void writeBinaryRecurs(std::ostream& s, const OcTreeNode& node) const
  {
    unsigned bits = 0;
    for (unsigned i = 0; i < 8; ++i)
    {
      if (!((node.mask_ >> i) & 1u))
        continue;
      const OcTreeNode& c = child(node, i);
      if (c.kids_ != 0)
        bits |= 3u << (2 * i);
      else if (isNodeOccupied(c))
        bits |= 2u << (2 * i);
      else
        bits |= 1u << (2 * i);
    }
    const char b[2] = { static_cast<char>(bits & 0xff), static_cast<char>((bits >> 8) & 0xff) };
    s.write(b, 2);
    for (unsigned i = 0; i < 8; ++i)
      if (((node.mask_ >> i) & 1u) && child(node, i).kids_ != 0)
        writeBinaryRecurs(s, child(node, i));
  }
  bool readFullRecurs(std::istream& s, OcTreeNode& node_in, unsigned depth)
  {
    float v;
    char m;
    s.read(reinterpret_cast<char*>(&v), 4);
    if (s.gcount() != 4)
      return false;
    s.read(&m, 1);
    if (s.gcount() != 1)
      return false;
    OcTreeNode* node = &node_in;
    const bool node_is_root = (node == &root_);
    const std::size_t self = node_is_root ? 0 : static_cast<std::size_t>(node - nodes_.data());
    node->value_ = v;
    const unsigned mask = static_cast<unsigned char>(m);
    if (mask && depth >= kDepth)
      return false;
    for (unsigned i = 0; i < 8; ++i)
    {
      if (!((mask >> i) & 1u))
        continue;
      OcTreeNode* me = node_is_root ? &root_ : &nodes_[self];
      OcTreeNode& c = childOrCreate(me, i);
      c = OcTreeNode();
      const std::size_t ci = static_cast<std::size_t>(&c - nodes_.data());
      if (!readFullRecurs(s, nodes_[ci], depth + 1))
        return false;
    }
    return true;
  }
  void writeFullRecurs(std::ostream& s, const OcTreeNode& node) const
  {
    const float v = node.value_;
    s.write(reinterpret_cast<const char*>(&v), 4);
    const char m = static_cast<char>(node.kids_ ? node.mask_ : 0);
    s.write(&m, 1);
    if (node.kids_)
      for (unsigned i = 0; i < 8; ++i)
        if ((node.mask_ >> i) & 1u)
          writeFullRecurs(s, child(node, i));
  }
  void updateInnerRecurs(OcTreeNode& node)
  {
    if (node.kids_ == 0)
      return;
    float best = -std::numeric_limits<float>::max();
    for (unsigned i = 0; i < 8; ++i)
    {
      if (!((node.mask_ >> i) & 1u))
        continue;
      OcTreeNode& c = child(node, i);
      updateInnerRecurs(c);
      if (c.value_ > best)
        best = c.value_;
    }
    node.value_ = best;
  }
  std::size_t countRecurs(const OcTreeNode& node) const
  {
    std::size_t n = 1;
    if (node.kids_)
      for (unsigned i = 0; i < 8; ++i)
        if ((node.mask_ >> i) & 1u)
          n += countRecurs(child(node, i));
    return n;
  }
  void ensureBounds()
  {
    if (bounds_valid_)
      return;
    const double big = 1e6;
    for (int a = 0; a < 3; ++a)
    {
      min_[a] = big;
      max_[a] = -big;
    }
    bool any = false;
    for (leaf_iterator it = begin_leafs(), e = end_leafs(); it != e; ++it)
    {
      any = true;
      const double size = it.getSize();
      const double half = size / 2.0;
      double c[3] = { it.getX() - half, it.getY() - half, it.getZ() - half };
      for (int a = 0; a < 3; ++a)
      {
        if (c[a] < min_[a])
          min_[a] = c[a];
        c[a] += size;
        if (c[a] > max_[a])
          max_[a] = c[a];
      }
    }
    if (!any)
      for (int a = 0; a < 3; ++a)
        min_[a] = max_[a] = 0.0;
    bounds_valid_ = true;
  }

  double resolution_;
  float occ_thres_log_{ 0.f };
  float clamp_min_{ -2.f };
  float clamp_max_{ 3.5f };
  OcTreeNode root_;
  bool has_root_{ false };
  std::vector<OcTreeNode> nodes_;
  std::size_t tree_size_{ 0 };
  bool size_valid_{ true };
  bool bounds_valid_{ false };
  double min_[3]{ 0, 0, 0 };
  double max_[3]{ 0, 0, 0 };
};

inline AbstractOcTree* AbstractOcTree::read(const std::string& filename)
{
  std::ifstream f(filename.c_str(), std::ios::in | std::ios::binary);
  if (!f.is_open())
    return nullptr;
  std::string line;
  if (!std::getline(f, line) || line.compare(0, 21, "# Octomap OcTree file") != 0)
    return nullptr;
  std::string id;
  uint64_t n = 0;
  double res = 0;
  if (!OcTree::readHeaderFields(f, id, n, res))
    return nullptr;
  if (id != "OcTree")
    return nullptr;  // other tree types (ColorOcTree, ...) are not OcTree; amcl3d's dynamic_cast would fail too
  OcTree* tree = new OcTree(res);
  if (n > 0 && !tree->readData(f))
  {
    delete tree;
    return nullptr;
  }
  return tree;
}
}  // namespace octomap
