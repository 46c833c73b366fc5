// mini_gtest.hpp — the handful of GoogleTest macros the reference's test files use, so their bodies can be mirrored
// here without GoogleTest (which is not installed in this image).  ASSERT_FLOAT_EQ keeps gtest's meaning: both sides
// rounded to float, at most 4 ULPs apart.
#pragma once

#include <sys/wait.h>
#include <unistd.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <exception>
#include <functional>
#include <string>
#include <vector>

namespace mini_gtest {

struct Registry {
  struct Entry {
    std::string name;
    std::function<void()> body;
  };
  static std::vector<Entry>& tests() {
    static std::vector<Entry> t;
    return t;
  }
  static int& failures() {
    static int f = 0;
    return f;
  }
  static bool& current_failed() {
    static bool f = false;
    return f;
  }
};

struct Registrar {
  Registrar(const char* name, std::function<void()> body) { Registry::tests().push_back({name, std::move(body)}); }
};

class Test {
 public:
  virtual ~Test() = default;
  virtual void SetUp() {}
  virtual void TearDown() {}
  virtual void TestBody() = 0;
  void Run() {
    SetUp();
    if (!Registry::current_failed()) TestBody();
    TearDown();
  }
};

inline bool AlmostEqualFloats(double a, double b) {
  float const fa = static_cast<float>(a), fb = static_cast<float>(b);
  if (std::isnan(fa) || std::isnan(fb)) return false;
  if (fa == fb) return true;
  std::int32_t ia, ib;
  std::memcpy(&ia, &fa, 4);
  std::memcpy(&ib, &fb, 4);
  auto biased = [](std::int32_t v) -> std::int64_t { return v < 0 ? -static_cast<std::int64_t>(v & 0x7fffffff) : v; };
  std::int64_t const d = biased(ia) - biased(ib);
  return (d < 0 ? -d : d) <= 4;
}

inline void Fail(const char* file, int line, const std::string& msg) {
  std::fprintf(stderr, "%s:%d: Failure\n  %s\n", file, line, msg.c_str());
  Registry::current_failed() = true;
}

// runs `statement` in a forked child; passes when the child dies from a signal or a non-zero exit
template <class F>
bool Dies(F&& statement) {
  std::fflush(nullptr);
  pid_t const pid = fork();
  if (pid == 0) {
    if (!freopen("/dev/null", "w", stderr)) _exit(0);
    statement();
    _exit(0);
  }
  int status = 0;
  waitpid(pid, &status, 0);
  return WIFSIGNALED(status) || (WIFEXITED(status) && WEXITSTATUS(status) != 0);
}

inline int RunAll(const char* filter) {
  int ran = 0;
  for (auto& t : Registry::tests()) {
    if (filter && *filter && t.name.find(filter) == std::string::npos) continue;
    Registry::current_failed() = false;
    std::printf("[ RUN      ] %s\n", t.name.c_str());
    try {  // GoogleTest reports an exception escaping SetUp / the body as that test's failure and carries on
      t.body();
    } catch (std::exception const& e) {
      Fail(t.name.c_str(), 0, std::string("C++ exception with description \"") + e.what() + "\" thrown in the test body or SetUp");
    } catch (...) {
      Fail(t.name.c_str(), 0, "unknown C++ exception thrown in the test body or SetUp");
    }
    ++ran;
    if (Registry::current_failed()) {
      ++Registry::failures();
      std::printf("[  FAILED  ] %s\n", t.name.c_str());
    } else {
      std::printf("[       OK ] %s\n", t.name.c_str());
    }
  }
  std::printf("[==========] %d tests ran, %d failed\n", ran, Registry::failures());
  return Registry::failures() == 0 ? 0 : 1;
}

}  // namespace mini_gtest

namespace testing {
using Test = mini_gtest::Test;
inline void InitGoogleTest(int*, char**) {}
}  // namespace testing

#define MG_CAT_(a, b) a##b
#define MG_CAT(a, b) MG_CAT_(a, b)

#define TEST(suite, name)                                                                                     \
  static void MG_CAT(suite##_##name, _body)();                                                                \
  static mini_gtest::Registrar MG_CAT(suite##_##name, _reg)(#suite "." #name, MG_CAT(suite##_##name, _body)); \
  static void MG_CAT(suite##_##name, _body)()

#define TEST_F(fixture, name)                                                                  \
  class MG_CAT(fixture##_##name, _Test) : public fixture {                                     \
   public:                                                                                     \
    void TestBody() override;                                                                  \
  };                                                                                           \
  static mini_gtest::Registrar MG_CAT(fixture##_##name, _reg)(#fixture "." #name, [] {         \
    MG_CAT(fixture##_##name, _Test) t;                                                         \
    t.Run();                                                                                   \
  });                                                                                          \
  void MG_CAT(fixture##_##name, _Test)::TestBody()

#define ASSERT_FLOAT_EQ(a, b)                                                                                          \
  do {                                                                                                                 \
    double const mg_a = static_cast<double>(a), mg_b = static_cast<double>(b);                                         \
    if (!mini_gtest::AlmostEqualFloats(mg_a, mg_b)) {                                                                  \
      char mg_buf[256];                                                                                                \
      std::snprintf(mg_buf, sizeof(mg_buf), "ASSERT_FLOAT_EQ(%s, %s): %.9g vs %.9g", #a, #b, mg_a, mg_b);              \
      mini_gtest::Fail(__FILE__, __LINE__, mg_buf);                                                                    \
      return;                                                                                                          \
    }                                                                                                                  \
  } while (0)

#define ASSERT_NEAR(a, b, tol)                                                                                         \
  do {                                                                                                                 \
    double const mg_a = static_cast<double>(a), mg_b = static_cast<double>(b);                                         \
    if (!(std::fabs(mg_a - mg_b) <= (tol))) {                                                                          \
      char mg_buf[256];                                                                                                \
      std::snprintf(mg_buf, sizeof(mg_buf), "ASSERT_NEAR(%s, %s, %s): %.12g vs %.12g", #a, #b, #tol, mg_a, mg_b);      \
      mini_gtest::Fail(__FILE__, __LINE__, mg_buf);                                                                    \
      return;                                                                                                          \
    }                                                                                                                  \
  } while (0)

#define ASSERT_TRUE(cond)                                                  \
  do {                                                                     \
    if (!(cond)) {                                                         \
      mini_gtest::Fail(__FILE__, __LINE__, "ASSERT_TRUE(" #cond ")");      \
      return;                                                              \
    }                                                                      \
  } while (0)

#define ASSERT_EQ(a, b) ASSERT_TRUE((a) == (b))

#define EXPECT_DEATH(statement, regex)                                                                   \
  do {                                                                                                   \
    if (!mini_gtest::Dies([&] { (void)(statement); }))                                                   \
      mini_gtest::Fail(__FILE__, __LINE__, "EXPECT_DEATH(" #statement "): the statement did not die");   \
  } while (0)

#define ASSERT_THROW(statement, exception_type)                                                          \
  do {                                                                                                   \
    bool mg_thrown = false;                                                                              \
    try {                                                                                                \
      (void)(statement);                                                                                 \
    } catch (exception_type const&) {                                                                    \
      mg_thrown = true;                                                                                  \
    }                                                                                                    \
    if (!mg_thrown) {                                                                                    \
      mini_gtest::Fail(__FILE__, __LINE__, "ASSERT_THROW(" #statement ", " #exception_type ")");         \
      return;                                                                                            \
    }                                                                                                    \
  } while (0)

#define RUN_ALL_TESTS() mini_gtest::RunAll(argc > 1 ? argv[1] : "")
