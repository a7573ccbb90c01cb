#pragma once  // TEST STUB (syntax check only)
#include <string>
namespace actionlib {
template <class A>
class SimpleActionClient {
 public:
  SimpleActionClient(const std::string &, bool = true);
  template <class G>
  void sendGoal(const G &);
  bool waitForServer();
  bool waitForResult();
};
}  // namespace actionlib
