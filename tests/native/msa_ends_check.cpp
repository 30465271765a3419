// rattle_b200/csrc/msa_ends.hpp (the product's restatement of correct.cpp:32-92) on cases read from stdin:
//   n ncol, then n lines "row seq qual" ('.' = empty string); prints the cleaned "row seq qual" lines.
#include <iostream>
#include "../../rattle_b200/csrc/msa_ends.hpp"
int main() {
    int n, ncol;
    while (std::cin >> n >> ncol) {
        std::vector<Read> reads(n);
        std::vector<std::string> rows(n);
        for (int i = 0; i < n; ++i) {
            std::cin >> rows[i] >> reads[i].seq >> reads[i].quality;
            if (rows[i] == ".") rows[i].clear();
            if (reads[i].seq == ".") reads[i].seq.clear();
            if (reads[i].quality == ".") reads[i].quality.clear();
        }
        fix_msa_ends(reads, rows);
        for (int i = 0; i < n; ++i)
            std::cout << (rows[i].empty() ? "." : rows[i]) << ' ' << (reads[i].seq.empty() ? "." : reads[i].seq) << ' '
                      << (reads[i].quality.empty() ? "." : reads[i].quality) << '\n';
    }
    return 0;
}
