/* TEST INFRASTRUCTURE (oracle) — not product code.
 *
 * Minimal stand-in for the un-vendored submodule lib/LSDynaReader
 * (pinned at bd095e76ee5d0f603dbc86af098471f31344bfe4 in the reference's
 * .gitmodules; not served with /root/reference).  It declares only the
 * members the reference touches in Domain_d::CreateFromLSDyna
 * (src/common/Domain_d.C:1647-1699) so the reference's own CPU sources compile
 * unmodified into oracle/_ref/.  The constructor (oracle/stub/lsdyna_stub_reader.cpp)
 * reads *NODE / *ELEMENT_SOLID cards in the format visible in
 * examples/input/*.k so that src/explicit/main.C can load "File" decks in the
 * parity tests; since the real reader is absent, `.k` parsing itself stays
 * PARITY UNPINNED.
 */
#ifndef WF_ORACLE_LSDYNA_STUB_H
#define WF_ORACLE_LSDYNA_STUB_H
#include <vector>
namespace LS_Dyna {
struct ls_node {
  int m_id;
  double m_x[3];
};
struct ls_element {
  int m_id;
  std::vector<int> node;
};
class lsdynaReader {
 public:
  lsdynaReader() : m_elem_count(0) {}
  explicit lsdynaReader(const char *path);
  std::vector<ls_node> m_node;
  std::vector<ls_element> m_elem;
  int m_elem_count;
};
}  // namespace LS_Dyna
#endif
