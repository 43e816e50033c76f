// tiny test harness for the C++ QA executables (gtest is not in the image)
#pragma once
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <string>
#include <vector>

struct qa_registry {
    struct entry {
        std::string name;
        std::function<void()> fn;
    };
    static std::vector<entry>& tests()
    {
        static std::vector<entry> t;
        return t;
    }
    static int& failures()
    {
        static int f = 0;
        return f;
    }
};
struct qa_reg {
    qa_reg(const char* name, std::function<void()> fn) { qa_registry::tests().push_back({ name, std::move(fn) }); }
};
#define QA_TEST(suite, name)                                      \
    static void suite##_##name();                                 \
    static qa_reg reg_##suite##_##name(#suite "." #name, suite##_##name); \
    static void suite##_##name()
#define EXPECT_TRUE(c)                                                                  \
    do {                                                                                \
        if (!(c)) {                                                                     \
            std::printf("  EXPECT failed %s:%d: %s\n", __FILE__, __LINE__, #c);         \
            qa_registry::failures()++;                                                  \
        }                                                                               \
    } while (0)
#define EXPECT_EQ(a, b) EXPECT_TRUE((a) == (b))

inline int qa_main(int argc, char** argv)
{
    std::string filter = argc > 1 ? argv[1] : "";
    int ran = 0, failed = 0;
    for (auto& t : qa_registry::tests()) {
        if (!filter.empty() && t.name.find(filter) == std::string::npos)
            continue;
        int before = qa_registry::failures();
        try {
            t.fn();
        } catch (const std::exception& e) {
            std::printf("  exception: %s\n", e.what());
            qa_registry::failures()++;
        }
        bool ok = qa_registry::failures() == before;
        std::printf("[%s] %s\n", ok ? "  OK  " : "FAILED", t.name.c_str());
        std::fflush(stdout);
        ran++;
        failed += !ok;
    }
    std::printf("%d tests, %d failed\n", ran, failed);
    return failed ? 1 : 0;
}
