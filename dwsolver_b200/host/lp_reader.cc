// lp_reader.cc — CPLEX-LP subset reader of the host library.
//
// Replaces the reference's glp_read_lp calls (/root/reference/src/dw_subprob.c:105-111 for the block
// files, src/dw_solver.c:192-198 for the master file); GLPK is not available here.  Accepts what the
// reference's fixtures use: sections objective / constraints / bounds /
// general|integer|binary / end, senses <= >= = (and =< => < >), default column bounds [0, +inf),
// bound forms `l <= x <= u`, `x >= l`, `x <= u`, `x = v`, `x free`, `+-inf[inity]`, `\` comments.
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string_view>

#include "lp_model.h"

namespace dwh {
namespace {

enum TokKind { T_NUM, T_NAME, T_OP };
struct Tok {                      // `text` points into the file image: no allocation per token
    TokKind kind;
    double num = 0.0;
    std::string_view text;
};

struct TokSpan {                  // a section's tokens: a window into the token array, no copy
    const Tok *b = nullptr;
    size_t n = 0;
    size_t size() const { return n; }
    const Tok &operator[](size_t i) const { return b[i]; }
    const Tok *begin() const { return b; }
    const Tok *end() const { return b + n; }
};

// character classes of the CPLEX-LP lexer, one table lookup per character
struct CharClass {
    unsigned char start[256], body[256];
    CharClass()
    {
        for (int c = 0; c < 256; ++c) {
            start[c] = std::isalpha(c) || (c && std::strchr("!\"#$%&()/,;?@_`'{}|~", c)) ? 1 : 0;
            body[c] = start[c] || std::isdigit(c) || c == '.';
        }
    }
};
const CharClass kChars;
inline bool name_start(char c) { return kChars.start[(unsigned char)c]; }
inline bool name_char(char c) { return kChars.body[(unsigned char)c]; }
inline bool digit(char c) { return c >= '0' && c <= '9'; }

std::vector<Tok> tokenize(const std::string &text)
{
    std::vector<Tok> toks;
    toks.reserve(text.size() / 4);
    size_t pos = 0, n = text.size();
    while (pos < n) {
        const char c = text[pos];
        if (c == '\\') { while (pos < n && text[pos] != '\n') ++pos; continue; }     // comment
        if (std::isspace((unsigned char)c)) { ++pos; continue; }
        if (digit(c) || (c == '.' && pos + 1 < n && digit(text[pos + 1]))) {
            // number: digits [. digits] [e[+-]digits]  (an exponent needs at least one digit after it)
            size_t e = pos;
            while (e < n && digit(text[e])) ++e;
            bool plain_integer = true;
            if (e < n && text[e] == '.') { plain_integer = false; ++e; while (e < n && digit(text[e])) ++e; }
            if (e < n && (text[e] == 'e' || text[e] == 'E')) {
                size_t f = e + 1;
                if (f < n && (text[f] == '+' || text[f] == '-')) ++f;
                if (f < n && digit(text[f])) {
                    while (f < n && digit(text[f])) ++f;
                    e = f;
                    plain_integer = false;
                }
            }
            Tok t; t.kind = T_NUM; t.text = std::string_view(text.data() + pos, e - pos);
            if (plain_integer && e - pos <= 15) {            // exact in a double; the common case in LP files
                long long v = 0;
                for (size_t i = pos; i < e; ++i) v = v * 10 + (text[i] - '0');
                t.num = (double)v;
            } else {
                // strtod on a bounded copy (on the file image it would also swallow forms the grammar above does
                // not allow, e.g. hexadecimal)
                char buf[64];
                if (e - pos < sizeof buf) {
                    std::memcpy(buf, text.data() + pos, e - pos); buf[e - pos] = 0;
                    t.num = std::strtod(buf, nullptr);
                } else t.num = std::strtod(text.substr(pos, e - pos).c_str(), nullptr);
            }
            toks.push_back(t);
            pos = e;
            continue;
        }
        if (name_start(c)) {
            size_t e = pos + 1;
            while (e < n && name_char(text[e])) ++e;
            Tok t; t.kind = T_NAME; t.text = std::string_view(text.data() + pos, e - pos);
            toks.push_back(t);
            pos = e;
            continue;
        }
        Tok t; t.kind = T_OP;
        if ((c == '<' || c == '>' || c == '=') && pos + 1 < n && (text[pos + 1] == '=' || text[pos + 1] == '<' || text[pos + 1] == '>')
            && !(c == '=' && text[pos + 1] == '=')) {
            t.text = std::string_view(text.data() + pos, 2); pos += 2;
        } else if (c == '<' || c == '>' || c == '=' || c == '+' || c == '-' || c == ':') {
            t.text = std::string_view(text.data() + pos, 1); ++pos;
        } else {
            throw std::runtime_error("LP syntax error near: " + text.substr(pos, 30));
        }
        toks.push_back(t);
    }
    return toks;
}

std::string lower(std::string_view s)
{
    std::string r(s);
    for (auto &ch : r) ch = (char)std::tolower((unsigned char)ch);
    return r;
}

enum Section { S_NONE, S_OBJ, S_OBJMAX, S_ROWS, S_BOUNDS, S_INTS, S_BINS, S_END };

Section section_of(const std::string &w)
{
    static const std::unordered_map<std::string, Section> words = {
        {"minimize", S_OBJ}, {"minimise", S_OBJ}, {"minimum", S_OBJ}, {"min", S_OBJ},
        {"maximize", S_OBJMAX}, {"maximise", S_OBJMAX}, {"maximum", S_OBJMAX}, {"max", S_OBJMAX},
        {"st", S_ROWS}, {"s.t.", S_ROWS}, {"st.", S_ROWS},
        {"bounds", S_BOUNDS}, {"bound", S_BOUNDS},
        {"general", S_INTS}, {"generals", S_INTS}, {"gen", S_INTS},
        {"integer", S_INTS}, {"integers", S_INTS}, {"int", S_INTS},
        {"binary", S_BINS}, {"binaries", S_BINS}, {"bin", S_BINS},
        {"end", S_END}};
    auto it = words.find(w);
    return it == words.end() ? S_NONE : it->second;
}

bool is_op(const Tok &t, const char *s) { return t.kind == T_OP && t.text == s; }
bool is_sense(const Tok &t)
{
    return t.kind == T_OP && (t.text == "<=" || t.text == ">=" || t.text == "=<" || t.text == "=>" || t.text == "<" ||
                              t.text == ">" || t.text == "=");
}
int sense_dir(std::string_view s)      // -1: <=, +1: >=, 0: =
{
    if (s == "<=" || s == "=<" || s == "<") return -1;
    if (s == ">=" || s == "=>" || s == ">") return 1;
    return 0;
}

struct Parser {
    LpModel lp;
    int col(std::string_view name_view)
    {
        // one hash per lookup; short names stay in the small-string buffer
        auto ins = lp.col_index.try_emplace(std::string(name_view), lp.n());
        if (!ins.second) return ins.first->second;
        lp.col_names.push_back(ins.first->first);
        lp.obj.push_back(0.0);
        lp.col_lb.push_back(0.0);
        lp.col_ub.push_back(kInf);
        return lp.n() - 1;
    }
    // [+|-] [coef] name ... until a sense operator, the end, or the label of the next row
    size_t linear(const TokSpan &st, size_t pos, std::vector<std::pair<int, double>> &terms, double &constant)
    {
        double sign = 1.0, coef = 0.0;
        bool have_coef = false;
        while (pos < st.size()) {
            const Tok &t = st[pos];
            if (is_sense(t)) break;
            if (is_op(t, "+")) { ++pos; continue; }
            if (is_op(t, "-")) { sign = -sign; ++pos; continue; }
            if (t.kind == T_NUM) { coef = have_coef ? coef * t.num : t.num; have_coef = true; ++pos; continue; }
            if (t.kind == T_NAME) {
                if (pos + 1 < st.size() && is_op(st[pos + 1], ":")) break;
                terms.emplace_back(col(t.text), sign * (have_coef ? coef : 1.0));
                sign = 1.0; have_coef = false;
                ++pos;
                continue;
            }
            throw std::runtime_error("unexpected token in a linear expression: " + std::string(t.text));
        }
        constant = have_coef ? sign * coef : 0.0;
        return pos;
    }
    // optional signs, then a number or inf[inity]
    bool value(const TokSpan &st, size_t &pos, double &v)
    {
        size_t p = pos;
        double sgn = 1.0;
        while (p < st.size() && (is_op(st[p], "+") || is_op(st[p], "-"))) { if (st[p].text == "-") sgn = -sgn; ++p; }
        if (p < st.size() && st[p].kind == T_NUM) { v = sgn * st[p].num; pos = p + 1; return true; }
        if (p < st.size() && st[p].kind == T_NAME) {
            const std::string w = lower(st[p].text);
            if (w == "inf" || w == "infinity") { v = sgn * kInf; pos = p + 1; return true; }
        }
        return false;
    }
    void rows(const TokSpan &st)
    {
        size_t pos = 0;
        while (pos < st.size()) {
            std::string name;
            if (st[pos].kind == T_NAME && pos + 1 < st.size() && is_op(st[pos + 1], ":")) { name = std::string(st[pos].text); pos += 2; }
            std::vector<std::pair<int, double>> terms;
            double constant = 0.0;
            pos = linear(st, pos, terms, constant);
            if (pos >= st.size() || !is_sense(st[pos])) throw std::runtime_error("constraint without sense");
            const int dir = sense_dir(st[pos].text);
            ++pos;
            double rhs = 0.0;
            if (!value(st, pos, rhs)) throw std::runtime_error("constraint without a numeric right-hand side");
            rhs -= constant;
            const int i = lp.m();
            lp.row_names.push_back(name.empty() ? "r." + std::to_string(i + 1) : name);
            lp.row_lb.push_back(dir <= 0 && dir != 0 ? -kInf : rhs);
            lp.row_ub.push_back(dir >= 0 && dir != 0 ? kInf : rhs);
            if (dir == 0) { lp.row_lb.back() = rhs; lp.row_ub.back() = rhs; }
            for (auto &t : terms) { lp.a_row.push_back(i); lp.a_col.push_back(t.first); lp.a_val.push_back(t.second); }
        }
    }
    void bounds(const TokSpan &st)
    {
        size_t pos = 0;
        while (pos < st.size()) {
            double v1 = 0.0;
            size_t p1 = pos;
            if (value(st, p1, v1)) {                         // "l <= x [<= u]"  or  "u >= x [>= l]"
                if (p1 + 1 >= st.size() || !is_sense(st[p1]) || st[p1 + 1].kind != T_NAME)
                    throw std::runtime_error("bad bounds entry");
                const int d1 = sense_dir(st[p1].text);
                const int j = col(st[p1 + 1].text);
                pos = p1 + 2;
                if (d1 < 0) lp.col_lb[j] = v1; else if (d1 > 0) lp.col_ub[j] = v1; else { lp.col_lb[j] = lp.col_ub[j] = v1; }
                if (pos < st.size() && is_sense(st[pos]) && sense_dir(st[pos].text) != 0) {
                    size_t p2 = pos + 1;
                    double v2 = 0.0;
                    if (value(st, p2, v2)) {
                        if (sense_dir(st[pos].text) < 0) lp.col_ub[j] = v2; else lp.col_lb[j] = v2;
                        pos = p2;
                    }
                }
                continue;
            }
            if (st[pos].kind != T_NAME) throw std::runtime_error("bad bounds entry at: " + std::string(st[pos].text));
            const int j = col(st[pos].text);
            ++pos;
            if (pos < st.size() && st[pos].kind == T_NAME && lower(st[pos].text) == "free") {
                lp.col_lb[j] = -kInf; lp.col_ub[j] = kInf;
                ++pos;
                continue;
            }
            if (pos >= st.size() || !is_sense(st[pos])) throw std::runtime_error("bad bounds entry for " + lp.col_names[j]);
            const int d = sense_dir(st[pos].text);
            ++pos;
            double v = 0.0;
            if (!value(st, pos, v)) throw std::runtime_error("bound without a value for " + lp.col_names[j]);
            if (d < 0) lp.col_ub[j] = v; else if (d > 0) lp.col_lb[j] = v; else { lp.col_lb[j] = lp.col_ub[j] = v; }
        }
    }
};

}  // namespace

LpModel parse_lp_text(const std::string &text)
{
    const std::vector<Tok> toks = tokenize(text);
    // split into sections ("subject to" / "such that" are two words; a keyword followed by ':' is a label)
    std::vector<std::pair<Section, TokSpan>> secs;
    Section cur = S_NONE;
    size_t sec_begin = 0;
    auto close = [&](size_t end) {
        if (cur != S_NONE) { TokSpan sp; sp.b = toks.data() + sec_begin; sp.n = end - sec_begin; secs.emplace_back(cur, sp); }
    };
    for (size_t i = 0; i < toks.size();) {
        Section sec = S_NONE;
        size_t adv = 1;
        // every keyword is at most 9 characters long and starts with one of m s b g i e (any case): almost every
        // name token of a large model is dismissed by this test without building a lower-case copy
        // ... and, like glp_read_lp, only a word that BEGINS A LINE can be a section keyword: a column that happens to be
        // called st, end, bin, int, ... inside an objective or constraint line is a column
        auto begins_line = [&](const Tok &t) {
            const char *p0 = text.data(), *p = t.text.data();
            while (p > p0 && (p[-1] == ' ' || p[-1] == '\t' || p[-1] == '\r')) --p;
            return p == p0 || p[-1] == '\n';
        };
        if (toks[i].kind == T_NAME && toks[i].text.size() <= 9 && std::strchr("mMsSbBgGiIeE", toks[i].text[0]) &&
            !(i + 1 < toks.size() && is_op(toks[i + 1], ":")) && begins_line(toks[i])) {
            const std::string w = lower(toks[i].text);
            if ((w == "subject" || w == "such") && i + 1 < toks.size() && toks[i + 1].kind == T_NAME &&
                (lower(toks[i + 1].text) == "to" || lower(toks[i + 1].text) == "that")) { sec = S_ROWS; adv = 2; }
            else sec = section_of(w);
        }
        if (sec != S_NONE) {
            close(i);
            cur = sec;
            i += adv;
            sec_begin = i;
            if (sec == S_END) { cur = S_NONE; break; }
            continue;
        }
        ++i;
        if (i == toks.size()) close(i);
    }

    Parser P;
    P.lp.col_index.reserve(toks.size() / 4);          // about one new name per four tokens in a wide model
    for (auto &sc : secs) {
        const TokSpan &st = sc.second;
        switch (sc.first) {
        case S_OBJ:
        case S_OBJMAX: {
            P.lp.maximize = sc.first == S_OBJMAX;
            size_t pos = 0;
            if (st.size() >= 2 && st[0].kind == T_NAME && is_op(st[1], ":")) pos = 2;
            std::vector<std::pair<int, double>> terms;
            double constant = 0.0;
            P.linear(st, pos, terms, constant);
            for (auto &t : terms) P.lp.obj[t.first] += t.second;
            P.lp.obj_const = constant;
            break;
        }
        case S_ROWS: P.rows(st); break;
        case S_BOUNDS: P.bounds(st); break;
        case S_INTS:
        case S_BINS:
            for (auto &t : st)
                if (t.kind == T_NAME) {
                    const int j = P.col(t.text);
                    if (sc.first == S_BINS) { P.lp.col_ub[j] = 1.0; }
                }
            break;
        default: break;
        }
    }
    return std::move(P.lp);
}

LpModel read_lp_file(const std::string &path)
{
    std::FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open LP file " + path);
    std::string text;
    char chunk[1 << 16];
    size_t got;
    while ((got = std::fread(chunk, 1, sizeof chunk, f)) > 0) text.append(chunk, got);
    const bool bad = std::ferror(f) != 0;
    std::fclose(f);
    if (bad) throw std::runtime_error("cannot read LP file " + path);
    try {
        return parse_lp_text(text);
    } catch (const std::exception &e) {
        throw std::runtime_error(path + ": " + e.what());
    }
}

}  // namespace dwh
