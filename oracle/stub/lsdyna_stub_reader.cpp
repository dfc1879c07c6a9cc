// TEST INFRASTRUCTURE (oracle) — not product code.
//
// Constructor of the stand-in LS_Dyna::lsdynaReader (oracle/stub/inc/lsdynaReader.h).  The reference's reader is the
// un-served submodule lib/LSDynaReader @ bd095e76; this one follows the card layout visible in
// examples/input/tetra_cyl.k / cyl_hex.k and the older in-tree reader test/lsdynaReader.C:170-205:
//   *NODE            id (8 columns) + x y z (16 columns each)
//   *ELEMENT_SOLID   eid pid n1..n8 (8 columns each); tetrahedra repeat their last node to fill 8 slots
// Node ids become 0-based indices in order of appearance (Domain_d::CreateFromLSDyna copies `node[]` straight into
// m_elnod, Domain_d.C:1677-1686); a solid is cut at its first repeated node (DEVLOG 20250609).
#include <cstdlib>
#include <fstream>
#include <map>
#include <string>

#include "lsdynaReader.h"

namespace LS_Dyna {

static std::string col(const std::string &l, size_t pos, size_t w) { return pos < l.size() ? l.substr(pos, w) : std::string(); }

lsdynaReader::lsdynaReader(const char *path) : m_elem_count(0) {
  std::ifstream f(path);
  std::string line;
  int sect = 0;  // 1 nodes, 2 solids
  std::map<int, int> idx;
  while (std::getline(f, line)) {
    while (!line.empty() && (line.back() == '\r' || line.back() == '\n')) line.pop_back();
    if (line.empty() || line[0] == '$') continue;
    if (line[0] == '*') {
      sect = line == "*NODE" ? 1 : (line.rfind("*ELEMENT_SOLID", 0) == 0 ? 2 : 0);
      continue;
    }
    if (sect == 1) {
      ls_node n;
      n.m_id = atoi(col(line, 0, 8).c_str());
      for (int d = 0; d < 3; d++) n.m_x[d] = atof(col(line, 8 + 16 * d, 16).c_str());
      idx[n.m_id] = (int)m_node.size();
      m_node.push_back(n);
    } else if (sect == 2) {
      ls_element e;
      e.m_id = atoi(col(line, 0, 8).c_str());
      for (size_t c = 2; c < 10; c++) {
        std::string t = col(line, 8 * c, 8);
        if (t.find_first_not_of(" \t") == std::string::npos) break;
        int id = atoi(t.c_str());
        auto it = idx.find(id);
        int k = it == idx.end() ? id - 1 : it->second;
        bool rep = false;
        for (int q : e.node) rep = rep || q == k;
        if (rep) break;
        e.node.push_back(k);
      }
      if (!e.node.empty()) m_elem.push_back(e);
    }
  }
  m_elem_count = (int)m_elem.size();
}

}  // namespace LS_Dyna
