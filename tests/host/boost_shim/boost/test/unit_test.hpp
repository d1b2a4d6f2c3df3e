// Minimal stand-in for the subset of Boost.Test the reference's unit tests use (BOOST_AUTO_TEST_CASE, BOOST_CHECK*,
// BOOST_ERROR), so those tests can be compiled UNCHANGED against the B200 build's host mirror on a machine without
// Boost.  Test infrastructure only.  Cases named in the environment variable JPGENC_SKIP_CASES (comma separated) are
// skipped; exit status = number of failed checks (capped at 255).
#pragma once
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace jpgenc_boost_test {
struct Case { const char* name; void (*fn)(); };
inline std::vector<Case>& registry() { static std::vector<Case> r; return r; }
inline int& failures() { static int n = 0; return n; }
inline int& checks() { static int n = 0; return n; }
inline const char*& current() { static const char* c = ""; return c; }
struct Registrar { Registrar(const char* n, void (*f)()) { registry().push_back({n, f}); } };
inline void report(const char* file, int line, const std::string& msg) {
    ++failures();
    std::cerr << file << "(" << line << "): error in \"" << current() << "\": " << msg << std::endl;
}
template <class T>
auto show(std::ostream& o, const T& v, int) -> decltype(o << +v, void()) { o << +v; }
template <class T>
auto show(std::ostream& o, const T& v, long) -> decltype(o << v, void()) { o << v; }
template <class T>
void show(std::ostream& o, const T&, ...) { o << "<value>"; }
template <class A, class B>
inline void check_equal(const A& a, const B& b, bool want_equal, const char* ea, const char* eb, const char* file, int line) {
    ++checks();
    if ((a == b) != want_equal) {
        std::ostringstream s;
        s << "check " << ea << (want_equal ? " == " : " != ") << eb << " failed [";
        show(s, a, 0);
        s << (want_equal ? " != " : " == ");
        show(s, b, 0);
        s << "]";
        report(file, line, s.str());
    }
}
template <class I, class J>
inline void check_collections(I a, I ae, J b, J be, const char* file, int line) {
    ++checks();
    std::size_t i = 0;
    bool ok = true;
    for (; a != ae && b != be; ++a, ++b, ++i)
        if (!(*a == *b)) {
            std::ostringstream s;
            s << "collections differ at position " << i << ": ";
            show(s, *a, 0);
            s << " != ";
            show(s, *b, 0);
            report(file, line, s.str());
            ok = false;
        }
    if (ok && (a != ae || b != be)) report(file, line, "collections differ in size");
}
inline int run_all() {
    std::string skip = std::getenv("JPGENC_SKIP_CASES") ? std::getenv("JPGENC_SKIP_CASES") : "";
    skip = "," + skip + ",";
    int ran = 0;
    for (const Case& c : registry()) {
        if (skip.find(std::string(",") + c.name + ",") != std::string::npos) { std::cout << "skipped " << c.name << std::endl; continue; }
        current() = c.name;
        const int before = failures();
        try { c.fn(); } catch (const std::exception& e) { report("?", 0, std::string("uncaught exception: ") + e.what()); }
        std::cout << (failures() == before ? "ok      " : "FAILED  ") << c.name << std::endl;
        ++ran;
    }
    std::cout << ran << " cases, " << checks() << " checks, " << failures() << " failures" << std::endl;
    return failures() > 255 ? 255 : failures();
}
}  // namespace jpgenc_boost_test

#define BOOST_AUTO_TEST_CASE(name)                                                      \
    static void name##_case();                                                          \
    static ::jpgenc_boost_test::Registrar name##_registrar(#name, &name##_case);        \
    static void name##_case()
#define BOOST_CHECK(expr)                                                               \
    do { ++::jpgenc_boost_test::checks();                                               \
         if (!(expr)) ::jpgenc_boost_test::report(__FILE__, __LINE__, "check " #expr " failed"); } while (0)
#define BOOST_CHECK_EQUAL(a, b) ::jpgenc_boost_test::check_equal((a), (b), true, #a, #b, __FILE__, __LINE__)
#define BOOST_CHECK_NE(a, b) ::jpgenc_boost_test::check_equal((a), (b), false, #a, #b, __FILE__, __LINE__)
#define BOOST_CHECK_EQUAL_COLLECTIONS(a, ae, b, be) ::jpgenc_boost_test::check_collections((a), (ae), (b), (be), __FILE__, __LINE__)
#define BOOST_ERROR(msg) ::jpgenc_boost_test::report(__FILE__, __LINE__, (msg))
