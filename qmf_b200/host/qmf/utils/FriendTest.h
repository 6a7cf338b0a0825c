// The reference's headers befriend its gtests (FRIEND_TEST from <gtest/gtest_prod.h>) so that they can reach
// private members; the same declarations here let those tests compile against this tree unmodified
// (make -C qmf_b200/host reftests).  gtest itself is not a dependency: the macro is declared when absent.
#pragma once
#ifndef FRIEND_TEST
#define FRIEND_TEST(test_case_name, test_name) friend class test_case_name##_##test_name##_Test
#endif
