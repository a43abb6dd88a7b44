// c2a_front.cpp — circom-subset front end + AST walker: host C++ restatement of
//   src/program.rs::compile (:18-74), src/process.rs (:22-764) and the parts of src/runtime.rs (:56-793) the walk needs.
// The reference parses with the iden3/circom crates (@ e8e125e, not in its tree); this file parses the subset the
// reference accepts (README.md:16-40) with a hand-written lexer / recursive-descent parser that applies the same
// desugaring the circom parser does (for -> while, `x++` / `x += e` -> `x = x + ...`, `a ==> b` -> `b <== a`,
// declarations with initialisers -> declaration + substitution).  The contract is the ORDER of add_signal / add_gate /
// add_connection calls and the signal ids / names they carry.
//
// Runtime model.  The reference deep-clones the whole Context for every `while` iteration / `if` body and merges
// variables and components back on exit (runtime.rs:151-187).  That is observationally a lexical scope: items declared
// inside are dropped, writes to items that already existed persist (a re-declared variable overwrites the outer one,
// :175-178), the return variable is always carried out (:180-184).  Here: one item STACK per call frame plus a stack of
// scope marks - push is O(1), pop truncates.  Template / function calls start an empty frame (runtime.rs:75-77).
// Temporaries (`random_<u32>` items, :229) are values, not entries; temporary SIGNALS still consume signal ids and are
// named "<ctx>.random_<id>" (the reference's suffix is a thread_rng draw, i.e. unspecified).
// The walk never touches a string: identifiers are interned after parsing (Symbols), `const_signal_<v>` items are keyed
// by their value, and signal names are kept as 16-byte records (one span record for a whole replayed instance) that are
// spelled out only when somebody asks for one (c2a_program_signal_name, the input / output prefix match, a host Compiler
// that wants the names).
// Recording.  The calls are recorded in the packed form the device emitter reads (Sink).  A (callable, arguments) pair is
// interpreted twice at most: a call runs in an empty context, so later instances are the first one's calls with shifted
// signal ids (Walker::handle_call).  With a host emitter attached they are replayed call by call; without one they are kept
// as c2a_replay records and expanded on the GPU (c2a_emit_compressed_device) or, on request, on the host (Sink::materialise).
// The walk runs on a thread with a 512 MB stack so that the call-depth guard fires before the native stack ends.
// u32 arithmetic on variables follows a release build: + * ** wrap, shifts use the low 5 bits, - / \ % error as in
// src/process.rs:649-750.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <new>
#include <pthread.h>
#include <memory>
#include <optional>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/c2a.h"

namespace front {

struct Error {
  int code;
  std::string text;
};
[[noreturn]] static void fail(int code, const std::string& text) { throw Error{code, text}; }
static void runtime_error(const std::string& what) { fail(C2A_PROG_RUNTIME_ERROR, "Runtime error: " + what); }

// ======================================================================================================= lexer
enum Tok { T_EOF, T_ID, T_NUM, T_STR, T_OP };
struct Token {
  Tok t;
  std::string s;
  int line;
};

static std::vector<Token> lex(const std::string& src) {
  static const char* ops[] = {"<<=", ">>=", "**=", "<==", "==>", "<--", "-->", "===", "&&", "||", "==", "!=", "<=", ">=", "<<", ">>", "**", "+=", "-=", "*=",
                              "/=", "\\=", "%=", "|=", "&=", "^=", "++", "--", nullptr};
  std::vector<Token> out;
  size_t i = 0, n = src.size();
  int line = 1;
  while (i < n) {
    char c = src[i];
    if (c == '\n') { ++line; ++i; continue; }
    if (isspace((unsigned char)c)) { ++i; continue; }
    if (c == '/' && i + 1 < n && src[i + 1] == '/') { while (i < n && src[i] != '\n') ++i; continue; }
    if (c == '/' && i + 1 < n && src[i + 1] == '*') {
      i += 2;
      while (i + 1 < n && !(src[i] == '*' && src[i + 1] == '/')) { if (src[i] == '\n') ++line; ++i; }
      i += 2;
      continue;
    }
    if (isalpha((unsigned char)c) || c == '_' || c == '$') {
      size_t j = i;
      while (j < n && (isalnum((unsigned char)src[j]) || src[j] == '_' || src[j] == '$')) ++j;
      out.push_back({T_ID, src.substr(i, j - i), line});
      i = j;
      continue;
    }
    if (isdigit((unsigned char)c)) {
      size_t j = i;
      if (c == '0' && j + 1 < n && (src[j + 1] == 'x' || src[j + 1] == 'X')) { j += 2; while (j < n && isxdigit((unsigned char)src[j])) ++j; }
      else while (j < n && isdigit((unsigned char)src[j])) ++j;
      out.push_back({T_NUM, src.substr(i, j - i), line});
      i = j;
      continue;
    }
    if (c == '"') {
      size_t j = i + 1;
      while (j < n && src[j] != '"') { if (src[j] == '\\') ++j; ++j; }
      out.push_back({T_STR, src.substr(i + 1, j - i - 1), line});
      i = j + 1;
      continue;
    }
    bool matched = false;
    for (int k = 0; ops[k]; ++k) {
      size_t L = strlen(ops[k]);
      if (src.compare(i, L, ops[k]) == 0) { out.push_back({T_OP, ops[k], line}); i += L; matched = true; break; }
    }
    if (matched) continue;
    out.push_back({T_OP, std::string(1, c), line});
    ++i;
  }
  out.push_back({T_EOF, "", line});
  return out;
}

// ========================================================================================================= AST
// ExpressionInfixOpcode in the order of c2a_gate_type via src/a_gate_type.rs:30-55
enum Infix { I_Mul, I_Div, I_Add, I_Sub, I_Pow, I_IntDiv, I_Mod, I_ShiftL, I_ShiftR, I_LesserEq, I_GreaterEq, I_Lesser, I_Greater, I_Eq, I_NotEq, I_BoolOr,
             I_BoolAnd, I_BitOr, I_BitAnd, I_BitXor };
static const uint32_t kGateOf[] = {C2A_AMul, C2A_ADiv, C2A_AAdd, C2A_ASub, C2A_APow, C2A_AIntDiv, C2A_AMod, C2A_AShiftL, C2A_AShiftR, C2A_ALEq, C2A_AGEq,
                                   C2A_ALt, C2A_AGt, C2A_AEq, C2A_ANeq, C2A_ABoolOr, C2A_ABoolAnd, C2A_ABitOr, C2A_ABitAnd, C2A_AXor};
enum Prefix { P_Sub, P_BoolNot, P_Complement };

struct Expr;
struct Callable;
using ExprP = std::shared_ptr<Expr>;
using Key = uint64_t;  // interned item name (see Symbols): what the walker compares instead of strings
struct Access {
  bool component;    // .name  vs  [expr]
  std::string name;
  ExprP index;
  Key key = 0;       // component: interned name (filled by Symbols::annotate)
};
struct Expr {
  enum Kind { Number, Variable, InfixOp, PrefixOp, Call, Unsupported } kind = Unsupported;
  std::string text;            // Number: literal; Variable/Call: name
  std::vector<Access> access;  // Variable
  int op = 0;
  ExprP l, r;
  std::vector<ExprP> args;
  // filled by Symbols::annotate after parsing
  Key key = 0;                       // Variable: interned name; Call: the callee's name symbol
  const Callable* callee = nullptr;  // Call: resolved definition (null: undefined)
  uint32_t depth = 1;                // height of the expression tree below this node (bounded by the parser: kMaxExprDepth)
  uint32_t num = 0;                  // Number: the literal's value ...
  bool num_overflow = false;         // ... or "does not fit u32" (src/process.rs:294-306; raised only when evaluated)
};
enum AssignOp { A_Var, A_Signal, A_ConstraintSignal };  // =  <--  <==
enum DataType { D_Variable, D_Signal, D_Component };
struct Stmt;
using StmtP = std::shared_ptr<Stmt>;
struct Stmt {
  enum Kind { Block, InitBlock, Substitution, Declaration, IfThenElse, While, Return, Assert, Unsupported } kind = Unsupported;
  std::vector<StmtP> stmts;    // Block / InitBlock
  std::string name;            // Substitution var / Declaration name
  std::vector<Access> access;  // Substitution
  AssignOp op = A_Var;
  ExprP e;                     // rhe / cond / value / arg
  DataType dtype = D_Variable;  // Declaration
  int sigkind = 0;             // 0 intermediate, 1 input, 2 output
  std::vector<ExprP> dims;
  StmtP a, b;                  // if / else / while body
  Key key = 0;                 // interned `name` (Symbols::annotate)
};
struct Callable {
  bool is_function = false;
  std::vector<std::string> params;
  std::vector<StmtP> body;
  std::vector<std::string> inputs, outputs;  // templates: declared input / output signal names
  std::vector<Key> param_keys, input_keys, output_keys;  // the same names interned (Symbols::annotate)
};
struct Program {
  std::map<std::string, Callable> defs;
  ExprP main;
};

// ====================================================================================================== parser
struct Parser {
  std::vector<Token> t;
  size_t p = 0;
  std::string file;
  explicit Parser(std::vector<Token> toks, std::string f) : t(std::move(toks)), file(std::move(f)) {}

  const Token& cur() const { return t[p]; }
  [[noreturn]] void err(const std::string& what) const {
    fail(C2A_PROG_PARSING_ERROR, "Parsing error: " + file + ":" + std::to_string(cur().line) + ": " + what + " near '" + cur().s + "'");
  }
  bool is_op(const char* s) const { return cur().t == T_OP && cur().s == s; }
  bool is_id(const char* s) const { return cur().t == T_ID && cur().s == s; }
  bool accept_op(const char* s) { if (is_op(s)) { ++p; return true; } return false; }
  bool accept_id(const char* s) { if (is_id(s)) { ++p; return true; } return false; }
  void expect_op(const char* s) { if (!accept_op(s)) err(std::string("expected '") + s + "'"); }
  std::string ident() { if (cur().t != T_ID) err("expected identifier"); return t[p++].s; }

  // ---- expressions: precedence of circom's grammar (lowest first): ?: || && cmp | ^ & shift +- */\% ** prefix
  // A library behind a C ABI must not take the caller down with it: expression NESTING (100 000 parentheses) would overflow the
  // parser's recursion, a left-deep CHAIN (a + a + ... two million times) the recursion of everything that later walks or frees
  // the tree.  Both are parse errors beyond kMaxExprDepth.
  static constexpr uint32_t kMaxExprDepth = 5000;
  uint32_t nest = 0;
  struct Nest {
    Parser& ps;
    explicit Nest(Parser& q) : ps(q) { if (++ps.nest > kMaxExprDepth) ps.err("expression nested too deeply"); }
    ~Nest() { --ps.nest; }
  };
  ExprP deep(ExprP e, uint32_t below) {
    e->depth = below + 1;
    if (e->depth > kMaxExprDepth) err("expression too deep");
    return e;
  }
  ExprP mk_infix(int op, ExprP l, ExprP r) {
    auto e = std::make_shared<Expr>(); e->kind = Expr::InfixOp; e->op = op; e->l = l; e->r = r;
    return deep(e, std::max(l->depth, r->depth));
  }
  ExprP expr() { Nest guard(*this); return ternary(); }
  ExprP ternary() {
    ExprP c = level(0);
    if (accept_op("?")) {  // InlineSwitchOp: parsed, rejected by the walker (src/process.rs:310)
      ExprP a = ternary();
      expect_op(":");
      ExprP b = ternary();
      auto e = std::make_shared<Expr>();
      e->kind = Expr::Unsupported;
      e->text = "InlineSwitchOp";
      return e;
    }
    return c;
  }
  ExprP level(int lv) {
    static const std::vector<std::vector<std::pair<const char*, int>>> L = {
        {{"||", I_BoolOr}}, {{"&&", I_BoolAnd}},
        {{"==", I_Eq}, {"!=", I_NotEq}, {"<=", I_LesserEq}, {">=", I_GreaterEq}, {"<", I_Lesser}, {">", I_Greater}},
        {{"|", I_BitOr}}, {{"^", I_BitXor}}, {{"&", I_BitAnd}}, {{"<<", I_ShiftL}, {">>", I_ShiftR}}, {{"+", I_Add}, {"-", I_Sub}},
        {{"*", I_Mul}, {"/", I_Div}, {"\\", I_IntDiv}, {"%", I_Mod}}, {{"**", I_Pow}}};
    if (lv == (int)L.size()) return prefix();
    ExprP l = level(lv + 1);
    while (true) {
      bool hit = false;
      for (auto& o : L[lv])
        if (is_op(o.first)) { ++p; l = mk_infix(o.second, l, level(lv + 1)); hit = true; break; }
      if (!hit) return l;
    }
  }
  ExprP prefix() {
    int op = -1;
    if (is_op("-")) op = P_Sub; else if (is_op("!")) op = P_BoolNot; else if (is_op("~")) op = P_Complement;
    if (op >= 0) {
      ++p;
      auto e = std::make_shared<Expr>();
      e->kind = Expr::PrefixOp;
      e->op = op;
      Nest guard(*this);
      e->r = prefix();
      return deep(e, e->r->depth);
    }
    return term();
  }
  std::vector<Access> accesses() {
    std::vector<Access> a;
    while (true) {
      if (accept_op("[")) { Access x{false, "", expr()}; expect_op("]"); a.push_back(x); }
      else if (is_op(".") ) { ++p; a.push_back(Access{true, ident(), nullptr}); }
      else return a;
    }
  }
  ExprP term() {
    auto e = std::make_shared<Expr>();
    if (cur().t == T_NUM) { e->kind = Expr::Number; e->text = t[p++].s; return e; }
    if (accept_op("(")) { ExprP in = expr(); expect_op(")"); return in; }
    if (accept_op("[")) {  // ArrayInLine: parsed, rejected by the walker
      if (!is_op("]")) { expr(); while (accept_op(",")) expr(); }
      expect_op("]");
      e->text = "ArrayInLine";
      return e;
    }
    if (cur().t == T_ID) {
      if (is_id("parallel")) { ++p; ExprP in = term(); (void)in; e->text = "ParallelOp"; return e; }
      std::string name = ident();
      if (accept_op("(")) {
        e->kind = Expr::Call;
        e->text = name;
        if (!is_op(")")) { e->args.push_back(expr()); while (accept_op(",")) e->args.push_back(expr()); }
        expect_op(")");
        for (auto& a : e->args) e->depth = std::max(e->depth, a->depth + 1);
        if (is_op("(")) {  // anonymous component T(..)(..)
          int depth = 0;
          do { if (is_op("(")) ++depth; if (is_op(")")) --depth; ++p; } while (depth > 0 && cur().t != T_EOF);
          e->kind = Expr::Unsupported;
          e->text = "AnonymousComp";
        }
        return e;
      }
      e->kind = Expr::Variable;
      e->text = name;
      e->access = accesses();
      for (auto& a : e->access) if (a.index) e->depth = std::max(e->depth, a.index->depth + 1);
      return e;
    }
    err("expected expression");
  }

  // ---- statements
  StmtP mk(Stmt::Kind k) { auto s = std::make_shared<Stmt>(); s->kind = k; return s; }
  StmtP block_of(std::vector<StmtP> v) { auto s = mk(Stmt::Block); s->stmts = std::move(v); return s; }
  StmtP subst(const std::string& var, std::vector<Access> acc, AssignOp op, ExprP rhe) {
    auto s = mk(Stmt::Substitution);
    s->name = var; s->access = std::move(acc); s->op = op; s->e = rhe;
    return s;
  }
  ExprP var_expr(const std::string& n, const std::vector<Access>& a) { auto e = std::make_shared<Expr>(); e->kind = Expr::Variable; e->text = n; e->access = a; return e; }
  ExprP num_expr(const std::string& v) { auto e = std::make_shared<Expr>(); e->kind = Expr::Number; e->text = v; return e; }

  // declaration list -> InitializationBlock [Declaration, (Substitution)]* in symbol order (circom: split_declaration_into_single_nodes)
  StmtP declaration() {
    DataType dt;
    int sigkind = 0;
    if (accept_id("var")) dt = D_Variable;
    else if (accept_id("component")) dt = D_Component;
    else if (accept_id("signal")) {
      dt = D_Signal;
      if (accept_id("input")) sigkind = 1; else if (accept_id("output")) sigkind = 2;
      if (accept_op("{")) { while (!accept_op("}")) { if (cur().t == T_EOF) err("unterminated tag list"); ++p; } }  // tags are ignored
    } else err("expected declaration");
    auto init = mk(Stmt::InitBlock);
    if (is_op("(")) err("tuple declarations are not supported");
    while (true) {
      auto d = mk(Stmt::Declaration);
      d->dtype = dt; d->sigkind = sigkind; d->name = ident();
      while (accept_op("[")) { d->dims.push_back(expr()); expect_op("]"); }
      init->stmts.push_back(d);
      AssignOp op;
      bool has = true;
      if (accept_op("=")) op = A_Var; else if (accept_op("<==")) op = A_ConstraintSignal; else if (accept_op("<--")) op = A_Signal; else has = false;
      if (has) init->stmts.push_back(subst(d->name, {}, op, expr()));
      if (!accept_op(",")) break;
    }
    return init;
  }

  // expression-statement: substitution in all its spellings
  StmtP simple_statement() {
    if (is_id("var") || is_id("signal") || is_id("component")) return declaration();
    ExprP lhs = expr();
    static const std::pair<const char*, int> compound[] = {{"+=", I_Add}, {"-=", I_Sub}, {"*=", I_Mul}, {"**=", I_Pow}, {"/=", I_Div}, {"\\=", I_IntDiv},
                                                           {"%=", I_Mod}, {"|=", I_BitOr}, {"&=", I_BitAnd}, {"^=", I_BitXor}, {"<<=", I_ShiftL}, {">>=", I_ShiftR}};
    auto need_var = [&](const ExprP& e) { if (e->kind != Expr::Variable) err("left-hand side must be a variable, signal or component access"); };
    if (accept_op("=")) { need_var(lhs); return subst(lhs->text, lhs->access, A_Var, expr()); }
    if (accept_op("<==")) { need_var(lhs); return subst(lhs->text, lhs->access, A_ConstraintSignal, expr()); }
    if (accept_op("<--")) { need_var(lhs); return subst(lhs->text, lhs->access, A_Signal, expr()); }
    if (accept_op("==>")) { ExprP r = expr(); need_var(r); return subst(r->text, r->access, A_ConstraintSignal, lhs); }
    if (accept_op("-->")) { ExprP r = expr(); need_var(r); return subst(r->text, r->access, A_Signal, lhs); }
    if (accept_op("===")) { expr(); auto s = mk(Stmt::Unsupported); s->name = "ConstraintEquality"; return s; }
    if (accept_op("++")) { need_var(lhs); return subst(lhs->text, lhs->access, A_Var, mk_infix(I_Add, var_expr(lhs->text, lhs->access), num_expr("1"))); }
    if (accept_op("--")) { need_var(lhs); return subst(lhs->text, lhs->access, A_Var, mk_infix(I_Sub, var_expr(lhs->text, lhs->access), num_expr("1"))); }
    for (auto& c : compound)
      if (accept_op(c.first)) { need_var(lhs); return subst(lhs->text, lhs->access, A_Var, mk_infix(c.second, var_expr(lhs->text, lhs->access), expr())); }
    err("expected an assignment");
  }

  StmtP statement() {
    if (accept_op("{")) {
      std::vector<StmtP> v;
      while (!accept_op("}")) { if (cur().t == T_EOF) err("unterminated block"); v.push_back(statement()); }
      return block_of(std::move(v));
    }
    if (accept_id("if")) {
      auto s = mk(Stmt::IfThenElse);
      expect_op("("); s->e = expr(); expect_op(")");
      s->a = statement();
      if (accept_id("else")) s->b = statement();
      return s;
    }
    if (accept_id("while")) {
      auto s = mk(Stmt::While);
      expect_op("("); s->e = expr(); expect_op(")");
      s->a = statement();
      return s;
    }
    if (accept_id("for")) {  // Block[init, While(cond, Block[body, step])]
      expect_op("(");
      StmtP init = simple_statement(); expect_op(";");
      ExprP cond = expr(); expect_op(";");
      StmtP step = simple_statement(); expect_op(")");
      StmtP body = statement();
      auto w = mk(Stmt::While);
      w->e = cond;
      w->a = block_of({body, step});
      return block_of({init, w});
    }
    if (accept_id("return")) { auto s = mk(Stmt::Return); s->e = expr(); expect_op(";"); return s; }
    if (accept_id("assert")) { auto s = mk(Stmt::Assert); expect_op("("); s->e = expr(); expect_op(")"); expect_op(";"); return s; }
    if (is_id("log")) {
      ++p; expect_op("(");
      int depth = 1;
      while (depth > 0) { if (cur().t == T_EOF) err("unterminated log"); if (is_op("(")) ++depth; if (is_op(")")) --depth; ++p; }
      expect_op(";");
      auto s = mk(Stmt::Unsupported); s->name = "LogCall"; return s;
    }
    StmtP s = simple_statement();
    expect_op(";");
    return s;
  }

  static void collect_io(const std::vector<StmtP>& body, Callable& c) {
    for (auto& s : body) {
      if (!s) continue;
      if (s->kind == Stmt::Declaration && s->dtype == D_Signal) {
        if (s->sigkind == 1) c.inputs.push_back(s->name);
        if (s->sigkind == 2) c.outputs.push_back(s->name);
      }
      collect_io(s->stmts, c);
      if (s->a) collect_io({s->a}, c);
      if (s->b) collect_io({s->b}, c);
    }
  }

  void definitions(Program& prog, const std::function<void(const std::string&)>& include) {
    while (cur().t != T_EOF) {
      if (accept_id("pragma")) { while (!accept_op(";")) { if (cur().t == T_EOF) err("unterminated pragma"); ++p; } continue; }
      if (accept_id("include")) { if (cur().t != T_STR) err("expected a file name"); std::string f = t[p++].s; expect_op(";"); include(f); continue; }
      if (is_id("template") || is_id("function")) {
        Callable c;
        c.is_function = cur().s == "function";
        ++p;
        while (is_id("custom") || is_id("parallel")) ++p;
        std::string name = ident();
        expect_op("(");
        if (!is_op(")")) { c.params.push_back(ident()); while (accept_op(",")) c.params.push_back(ident()); }
        expect_op(")");
        StmtP b = statement();
        if (b->kind != Stmt::Block) err("expected a body");
        c.body = b->stmts;
        if (!c.is_function) collect_io(c.body, c);
        prog.defs[name] = std::move(c);
        continue;
      }
      if (accept_id("component")) {
        if (!accept_id("main")) err("only `component main` is allowed at top level");
        if (accept_op("{")) { while (!accept_op("}")) { if (cur().t == T_EOF) err("unterminated public list"); ++p; } }
        expect_op("=");
        prog.main = expr();
        expect_op(";");
        continue;
      }
      err("expected template, function, include, pragma or component main");
    }
  }
};

static std::string read_file(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) fail(C2A_PROG_PARSING_ERROR, "Parsing error: cannot open " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}
static std::string dir_of(const std::string& path) {
  size_t k = path.find_last_of('/');
  return k == std::string::npos ? std::string(".") : path.substr(0, k);
}
static void parse_into(Program& prog, const std::string& src, const std::string& file, const std::string& dir, std::set<std::string>& seen) {
  Parser ps(lex(src), file);
  ps.definitions(prog, [&](const std::string& inc) {
    std::string path = inc.size() && inc[0] == '/' ? inc : dir + "/" + inc;
    if (path.size() < 7 || path.compare(path.size() - 7, 7, ".circom") != 0) path += ".circom";
    if (!seen.insert(path).second) return;
    parse_into(prog, read_file(path), path, dir_of(path), seen);
  });
}

// ===================================================================================================== symbols
// Item keys.  An identifier is its interned id; the items process.rs creates for constants ("const_signal_<value>",
// :558-579) are keyed by the value, and an identifier that happens to be spelled that way maps to the same key, so the two
// still meet in one namespace exactly as the reference's string-keyed map does.
constexpr Key kConstKey = 1ull << 32;
struct Symbols {
  std::unordered_map<std::string, uint32_t> ids;
  std::vector<std::string> strs;
  uint32_t intern(const std::string& s) {
    auto it = ids.find(s);
    if (it != ids.end()) return it->second;
    uint32_t id = (uint32_t)strs.size();
    strs.push_back(s);
    ids.emplace(s, id);
    return id;
  }
  Key key(const std::string& s) {
    static const char kPrefix[] = "const_signal_";
    const size_t L = sizeof(kPrefix) - 1;
    if (s.size() > L && s.size() <= L + 10 && s.compare(0, L, kPrefix) == 0) {
      bool digits = true;
      unsigned long long v = 0;
      for (size_t i = L; i < s.size(); ++i) { if (!isdigit((unsigned char)s[i])) { digits = false; break; } v = v * 10 + (s[i] - '0'); }
      if (digits && v <= 0xFFFFFFFFull && std::to_string(v) == s.substr(L)) return kConstKey | v;  // canonical spelling only
    }
    return intern(s);
  }
  std::string str(Key k) const { return k >= kConstKey ? "const_signal_" + std::to_string((uint32_t)k) : strs[(size_t)k]; }

  // ---- one pass over the parsed program: names -> keys, literals -> values, calls -> definitions
  void annotate(Program& prog) {
    for (auto& kv : prog.defs) {
      Callable& c = kv.second;
      c.param_keys.clear(); c.input_keys.clear(); c.output_keys.clear();
      for (auto& p : c.params) c.param_keys.push_back(key(p));
      for (auto& p : c.inputs) c.input_keys.push_back(key(p));
      for (auto& p : c.outputs) c.output_keys.push_back(key(p));
      for (auto& s : c.body) stmt(prog, s.get());
    }
    if (prog.main) expr(prog, prog.main.get());
  }
  void accesses(Program& prog, std::vector<Access>& acc) {
    for (auto& a : acc) {
      if (a.component) a.key = key(a.name);
      else if (a.index) expr(prog, a.index.get());
    }
  }
  void expr(Program& prog, Expr* e) {
    if (!e) return;
    switch (e->kind) {
      case Expr::Number: {
        const std::string& s = e->text;
        unsigned long long v = 0;
        bool hex = s.size() > 2 && (s[1] == 'x' || s[1] == 'X');
        e->num_overflow = false;
        for (size_t i = hex ? 2 : 0; i < s.size(); ++i) {
          int d = isdigit((unsigned char)s[i]) ? s[i] - '0' : (tolower(s[i]) - 'a' + 10);
          v = v * (hex ? 16 : 10) + d;
          if (v > 0xFFFFFFFFull) { e->num_overflow = true; break; }
        }
        e->num = (uint32_t)v;
        break;
      }
      case Expr::Variable: e->key = key(e->text); accesses(prog, e->access); break;
      case Expr::Call: {
        e->key = intern(e->text);
        auto def = prog.defs.find(e->text);
        e->callee = def == prog.defs.end() ? nullptr : &def->second;
        for (auto& a : e->args) expr(prog, a.get());
        break;
      }
      default: break;
    }
    expr(prog, e->l.get());
    expr(prog, e->r.get());
  }
  void stmt(Program& prog, Stmt* s) {
    if (!s) return;
    if (s->kind == Stmt::Substitution || s->kind == Stmt::Declaration) s->key = key(s->name);
    accesses(prog, s->access);
    expr(prog, s->e.get());
    for (auto& d : s->dims) expr(prog, d.get());
    for (auto& c : s->stmts) stmt(prog, c.get());
    stmt(prog, s->a.get());
    stmt(prog, s->b.get());
  }
};

// ===================================================================================================== runtime
// a few elements inline, the rest on the heap (access paths and dimension lists are almost always <= 4 long)
template <class T, int N>
struct Small {
  T a[N];
  uint32_t n = 0;
  std::vector<T> more;
  Small() = default;
  Small(const Small& o) : n(o.n), more(o.more) { for (uint32_t i = 0; i < n && i < (uint32_t)N; ++i) a[i] = o.a[i]; }
  Small(Small&& o) noexcept : n(o.n), more(std::move(o.more)) { for (uint32_t i = 0; i < n && i < (uint32_t)N; ++i) a[i] = o.a[i]; }
  Small& operator=(const Small& o) { n = o.n; more = o.more; for (uint32_t i = 0; i < n && i < (uint32_t)N; ++i) a[i] = o.a[i]; return *this; }
  Small& operator=(Small&& o) noexcept { n = o.n; more = std::move(o.more); for (uint32_t i = 0; i < n && i < (uint32_t)N; ++i) a[i] = o.a[i]; return *this; }
  void push_back(const T& v) { if (n < (uint32_t)N) a[n] = v; else more.push_back(v); ++n; }
  void pop_back() { if (n > (uint32_t)N) more.pop_back(); --n; }
  uint32_t size() const { return n; }
  bool empty() const { return n == 0; }
  const T& operator[](uint32_t i) const { return i < (uint32_t)N ? a[i] : more[i - N]; }
};
using Path = Small<uint32_t, 6>;

template <class T>
struct Nested {
  bool is_array = false;
  std::vector<Nested<T>> arr;
  T val{};
};
using SigTree = Nested<uint32_t>;
struct SigMap {  // a component's input / output signals by name (runtime.rs: HashMap<String, Signal>); a handful of entries
  std::vector<std::pair<Key, SigTree>> v;
  const SigTree* find(Key k) const {
    for (auto& e : v) if (e.first == k) return &e.second;
    return nullptr;
  }
  void set(Key k, const SigTree& t) {
    for (auto& e : v) if (e.first == k) { e.second = t; return; }
    v.emplace_back(k, t);
  }
};
struct Item {
  DataType type = D_Variable;
  Nested<std::optional<uint32_t>> var;
  SigTree sig;
  Nested<SigMap> comp;
};
struct SubAccess {
  bool component;
  Key v;  // component: the signal's name key; otherwise the index
};
using AccessList = Small<SubAccess, 4>;
// result of process_expression: a named item access, or a temporary
struct Ref {
  enum Kind { Named, TempVar, TempSignal, TempComp } kind = TempVar;
  Key name = 0;
  AccessList access;
  std::optional<uint32_t> value;  // TempVar
  uint32_t signal = 0;            // TempSignal
  std::shared_ptr<SigMap> comp;   // TempComp
};

template <class T>
static const Nested<T>& nested_get(const Nested<T>& v, const Path& path) {  // runtime.rs:668-688
  const Nested<T>* cur = &v;
  for (uint32_t k = 0; k < path.size(); ++k) {
    const uint32_t i = path[k];
    if (!cur->is_array) runtime_error("Access Error");
    if (i >= cur->arr.size()) runtime_error("Index out of bounds");
    cur = &cur->arr[i];
  }
  return *cur;
}
template <class T>
static Nested<T>& nested_get_mut(Nested<T>& v, const Path& path) {
  return const_cast<Nested<T>&>(nested_get(const_cast<const Nested<T>&>(v), path));
}
template <class T, class Leaf>
static Nested<T> nested_new(const Path& dims, uint32_t k, Leaf&& leaf) {
  Nested<T> n;
  if (k == dims.size()) { n.val = leaf(); return n; }
  n.is_array = true;
  n.arr.reserve(dims[k]);
  for (uint32_t i = 0; i < dims[k]; ++i) n.arr.push_back(nested_new<T>(dims, k + 1, leaf));
  return n;
}
static Path access_to_u32(const AccessList& a) {  // runtime.rs:701-710
  Path v;
  for (uint32_t k = 0; k < a.size(); ++k) { if (a[k].component) runtime_error("Access Error"); v.push_back((uint32_t)a[k].v); }
  return v;
}

// One call frame: the items in declaration order (a stack: a scope's items are the ones above its mark) and the open scopes.
// Lookups scan the stack from the top while it is short and go through an index once it is not.
struct Frame {
  uint32_t ctx = 0;  // name of the context = the callee's name symbol ("0" for the root, runtime.rs:63-68)
  std::vector<Key> keys;    // keys[i] names items[i]; kept apart so that a scan touches two cache lines, not one per item
  std::vector<Item> items;
  std::vector<uint32_t> marks;  // items.size() when each open scope began
  std::unordered_map<Key, uint32_t> index;
  bool indexed = false;
  static constexpr size_t kScan = 16;

  void reset(uint32_t ctx_sym) {
    ctx = ctx_sym;
    keys.clear();
    items.clear();
    marks.clear();
    marks.push_back(0);
    if (indexed) { index.clear(); indexed = false; }
  }
  size_t size() const { return keys.size(); }
  Item* find(Key k) {
    if (!indexed) {
      for (size_t i = keys.size(); i-- > 0;)
        if (keys[i] == k) return &items[i];
      return nullptr;
    }
    auto it = index.find(k);
    return it == index.end() ? nullptr : &items[it->second];
  }
  Item& push(Key k, Item&& item) {
    keys.push_back(k);
    items.push_back(std::move(item));
    if (indexed) index[k] = (uint32_t)keys.size() - 1;
    else if (keys.size() > kScan) {
      for (uint32_t i = 0; i < keys.size(); ++i) index[keys[i]] = i;
      indexed = true;
    }
    return items.back();
  }
  void truncate(size_t n) {
    if (indexed) for (size_t i = n; i < keys.size(); ++i) index.erase(keys[i]);
    keys.resize(n);
    items.resize(n);
  }
};

// Growable array of trivially copyable records on realloc(): for the blocks the recorded stream reaches (hundreds of MB), glibc
// grows the mapping in place (mremap) - no copy, and only the new pages are touched.  std::vector allocates, copies and
// first-touches a whole new block at every doubling, which tripled the cost of the replayed appends below.
template <class T>
struct PodVec {
  T* p = nullptr;
  size_t n = 0, cap = 0;
  PodVec() = default;
  PodVec(const PodVec&) = delete;
  PodVec& operator=(const PodVec&) = delete;
  PodVec(PodVec&& o) noexcept : p(o.p), n(o.n), cap(o.cap) { o.p = nullptr; o.n = o.cap = 0; }
  PodVec& operator=(PodVec&& o) noexcept { if (this != &o) { free(p); p = o.p; n = o.n; cap = o.cap; o.p = nullptr; o.n = o.cap = 0; } return *this; }
  ~PodVec() { free(p); }
  size_t size() const { return n; }
  size_t capacity() const { return cap; }
  T* data() { return p; }
  const T* data() const { return p; }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
  void reserve(size_t m) {
    if (m <= cap) return;
    size_t c = std::max<size_t>(m, std::max<size_t>(2 * cap, 1024));
    T* q = (T*)realloc(p, c * sizeof(T));
    if (!q) throw std::bad_alloc();
    p = q;
    cap = c;
  }
  void push_back(const T& v) { if (n == cap) reserve(n + 1); p[n++] = v; }
  void resize(size_t m) { reserve(m); if (m > n) memset((void*)(p + n), 0, (m - n) * sizeof(T)); n = m; }
  void resize_uninit(size_t m) { reserve(m); n = m; }
  void append_self(size_t b, size_t e) {  // append a copy of [b, e) of this array
    if (e <= b) return;
    reserve(n + (e - b));
    memcpy((void*)(p + n), (const void*)(p + b), (e - b) * sizeof(T));
    n += e - b;
  }
};

struct SigName {  // "<ctx>.<base>[i][j]..." spelled on demand (runtime.rs:594-608, process.rs:464-474, 558-579)
  uint32_t ctx;       // context name symbol
  uint32_t kind_n;    // kind (2 bits: 0 declared name, 1 const_signal_<a>, 2 random_<a>) | number of indices << 2
  uint32_t a;         // name symbol / constant value (random_<id>: the suffix is the signal id itself)
  uint32_t idx_off;   // first index in Sink::idx
};

// Where add_signal / add_gate / add_connection go.  The calls are recorded directly in the PACKED form the device emitter reads
// (include/c2a.h: one kind byte per call, u32 payload words for gates and connections; the walker numbers its signals 0, 1, 2,
// ... as it declares them, src/runtime.rs:120-125, so a signal call has no payload): 6 B per call instead of a 16-byte record,
// nothing to convert before the H2D copy.  Constant values and the name records travel beside the stream.
struct Sink {
  c2a_compiler* into = nullptr;
  PodVec<uint8_t> kinds;         // per call: kind | gate type << 2
  PodVec<uint32_t> words;        // gate: lhs, rhs, out; connection: a, b (signal ids)
  PodVec<uint32_t> const_ids;    // the constant signals in declaration order ...
  PodVec<uint32_t> const_vals;   // ... and their values
  PodVec<SigName> names;         // by signal id; the records of a REPLAYED id range are never written (nor their pages touched):
  struct Span { uint32_t dst, len, src; };  // ... ids [dst, dst + len) are named like [src, src + len), src < dst
  std::vector<Span> spans;       // ascending dst, disjoint
  PodVec<uint32_t> idx;          // array indices of the declared names
  PodVec<c2a_event> aos;         // the same calls as c2a_event records, written when somebody asks (c2a_program_events)
  bool aos_valid = false;
  // Without a host emitter attached, a replayed instance is not even copied: kinds / words grow by its length, the range stays
  // unwritten (untouched pages) and a c2a_replay record says where it comes from.  The device expands the records itself
  // (c2a_emit_compressed_device: the host ships every distinct instance once); materialise() carries them out on the host for
  // callers that want the full arrays.  gen = 1 + the largest gen among the records inside the source range (0 if none): all
  // records of one gen can be expanded concurrently once the smaller gens are done.
  std::vector<c2a_replay> replays;  // ascending dst
  size_t replays_done = 0;
  uint32_t max_gen = 0;
  bool lazy = false;
  Symbols sym;
  struct Mark { uint64_t k, w, c; uint32_t id; size_t r; };  // a position in the recording
  Mark mark() const { return Mark{kinds.size(), words.size(), const_ids.size(), (uint32_t)names.size(), replays.size()}; }
  uint32_t max_gen_between(const Mark& b, const Mark& e) const {
    uint32_t g = 0;
    for (size_t i = b.r; i < e.r; ++i) g = std::max(g, replays[i].gen);
    return g;
  }
  void materialise() {
    for (; replays_done < replays.size(); ++replays_done) {
      const c2a_replay& r = replays[replays_done];
      if (r.k_len) memcpy(kinds.data() + r.k_dst, kinds.data() + r.k_src, r.k_len);
      const uint32_t* src = words.data() + r.w_src;
      uint32_t* dst = words.data() + r.w_dst;
      for (uint64_t i = 0; i < r.w_len; ++i) dst[i] = src[i] + r.delta;
    }
  }

  void check(int st) {
    if (st == C2A_OK) return;
    std::string why = st == C2A_ERR_INVALID_ARGUMENT || st == C2A_ERR_REFERENCE_PANIC ? c2a_compiler_last_error(into) : c2a_status_string(st);
    fail(C2A_PROG_CIRCUIT_ERROR, "Circuit error: " + why);
  }
  uint32_t name_source(uint32_t id) const {  // the id whose record names this one (itself unless it lies in replayed ranges)
    while (!spans.empty()) {
      size_t lo = 0, hi = spans.size();  // last span with dst <= id
      while (lo < hi) { size_t mid = (lo + hi) / 2; if (spans[mid].dst <= id) lo = mid + 1; else hi = mid; }
      if (lo == 0) break;
      const Span& sp = spans[lo - 1];
      if (id - sp.dst >= sp.len) break;
      id = sp.src + (id - sp.dst);
    }
    return id;
  }
  std::string name_of(uint32_t id) const {
    const SigName& n = names[name_source(id)];
    std::string s = sym.strs[n.ctx];
    s += '.';
    switch (n.kind_n & 3u) {
      case 0: s += sym.strs[n.a]; break;
      case 1: s += "const_signal_"; s += std::to_string(n.a); break;
      default: s += "random_"; s += std::to_string(id); break;  // the suffix is the signal's own id
    }
    for (uint32_t k = 0; k < (n.kind_n >> 2); ++k) { s += '['; s += std::to_string(idx[n.idx_off + k]); s += ']'; }
    return s;
  }
  void add_signal(uint32_t id, uint32_t ctx, Key name, const Path* indices, bool random, std::optional<uint32_t> value) {
    if (id != names.size()) fail(C2A_PROG_RUNTIME_ERROR, "Runtime error: internal: signal ids are not sequential");
    SigName n;
    n.ctx = ctx;
    n.idx_off = (uint32_t)idx.size();
    const uint32_t ni = indices ? indices->size() : 0u;
    if (random) { n.kind_n = 2u; n.a = 0; }
    else if (name >= kConstKey) { n.kind_n = 1u | (ni << 2); n.a = (uint32_t)name; }
    else { n.kind_n = 0u | (ni << 2); n.a = (uint32_t)name; }
    for (uint32_t k = 0; k < ni; ++k) idx.push_back((*indices)[k]);
    names.push_back(n);
    kinds.push_back(value ? (uint8_t)C2A_EV_SIGNAL_CONST : (uint8_t)C2A_EV_SIGNAL);
    if (value) { const_ids.push_back(id); const_vals.push_back(*value); }
    if (into) check(c2a_add_signal(into, id, name_of(id).c_str(), value.has_value(), value.value_or(0)));
  }
  void add_gate(uint32_t op, uint32_t l, uint32_t r, uint32_t o) {
    kinds.push_back((uint8_t)(C2A_EV_GATE | (op << 2)));
    words.reserve(words.size() + 3);
    words.push_back(l); words.push_back(r); words.push_back(o);
    if (into) check(c2a_add_gate(into, op, l, r, o));
  }
  void add_connection(uint32_t a, uint32_t b) {
    kinds.push_back((uint8_t)C2A_EV_CONNECT);
    words.push_back(a); words.push_back(b);
    if (into) check(c2a_add_connection(into, a, b));
  }
  // Replay of an earlier instance of the same callable with the same arguments (Walker::handle_call): the calls recorded
  // between the marks b and e are recorded again with every signal id moved by delta.  A call starts an empty context, so each
  // id in the slice belongs to the slice - every payload word is one, and the kind bytes and name records do not change.
  void replay(const Mark& b, const Mark& e, uint32_t delta, uint32_t src_gen) {
    const Mark at = mark();
    if (lazy) {
      kinds.resize_uninit(kinds.size() + (e.k - b.k));
      words.resize_uninit(words.size() + (e.w - b.w));
      replays.push_back(c2a_replay{at.k, b.k, e.k - b.k, at.w, b.w, e.w - b.w, delta, src_gen + 1});
      max_gen = std::max(max_gen, src_gen + 1);
    } else kinds.append_self(b.k, e.k);
    if (e.id > b.id) {               // the names: one span record instead of 16 bytes per signal
      names.resize_uninit(names.size() + (e.id - b.id));
      spans.push_back(Span{at.id, e.id - b.id, b.id});
    }
    const_vals.append_self(b.c, e.c);
    const_ids.reserve(const_ids.size() + (e.c - b.c));
    for (uint64_t i = b.c; i < e.c; ++i) const_ids.push_back(const_ids[i] + delta);
    if (!lazy) {
      words.reserve(words.size() + (e.w - b.w));
      const uint32_t* src = words.data() + b.w;
      uint32_t* dst = words.data() + words.size();
      const uint64_t n = e.w - b.w;
      for (uint64_t i = 0; i < n; ++i) dst[i] = src[i] + delta;
      words.n += n;
    }
    if (into) {  // a host emitter is attached: it sees the calls one by one, as if they had been interpreted
      uint64_t w = at.w, c = at.c;
      uint32_t id = at.id;
      for (uint64_t i = at.k; i < kinds.size(); ++i) {
        const uint32_t kb = kinds[i];
        switch (kb & 3u) {
          case C2A_EV_SIGNAL: check(c2a_add_signal(into, id, name_of(id).c_str(), 0, 0)); ++id; break;
          case C2A_EV_SIGNAL_CONST: check(c2a_add_signal(into, id, name_of(id).c_str(), 1, const_vals[c++])); ++id; break;
          case C2A_EV_GATE: check(c2a_add_gate(into, kb >> 2, words[w], words[w + 1], words[w + 2])); w += 3; break;
          default: check(c2a_add_connection(into, words[w], words[w + 1])); w += 2; break;
        }
      }
    }
  }
  // the recording as c2a_event records (include/c2a.h), for callers of c2a_program_events
  const PodVec<c2a_event>& events() {
    if (aos_valid) return aos;
    materialise();
    aos.resize_uninit(kinds.size());
    uint64_t w = 0, c = 0;
    uint32_t id = 0;
    for (uint64_t i = 0; i < kinds.size(); ++i) {
      const uint32_t kb = kinds[i];
      switch (kb & 3u) {
        case C2A_EV_SIGNAL: aos[i] = c2a_event{(uint32_t)C2A_EV_SIGNAL, id++, 0, 0}; break;
        case C2A_EV_SIGNAL_CONST: aos[i] = c2a_event{(uint32_t)C2A_EV_SIGNAL_CONST, id++, const_vals[c++], 0}; break;
        case C2A_EV_GATE: aos[i] = c2a_event{(uint32_t)C2A_EV_GATE | ((kb >> 2) << 8), words[w], words[w + 1], words[w + 2]}; w += 3; break;
        default: aos[i] = c2a_event{(uint32_t)C2A_EV_CONNECT, words[w], words[w + 1], 0}; w += 2; break;
      }
    }
    aos_valid = true;
    return aos;
  }
};

struct Walker {
  const Program& prog;
  Sink& ac;
  std::vector<std::unique_ptr<Frame>> frames;  // [0, nframes) live; the rest are kept for reuse (their vectors stay allocated)
  size_t nframes = 0;
  uint32_t next_signal_id = 0;  // runtime.rs:120-125
  uint64_t depth = 0;
  const Key kReturn;            // "function_return_value"

  Walker(const Program& p, Sink& s) : prog(p), ac(s), kReturn(s.sym.key("function_return_value")) {
    push_frame(s.sym.intern("0"));  // runtime.rs:63-68: the root context is named "0"
  }
  Frame& ctx() { return *frames[nframes - 1]; }
  void push_frame(uint32_t ctx_sym) {
    if (nframes == frames.size()) frames.emplace_back(new Frame());
    frames[nframes++]->reset(ctx_sym);
  }
  void pop_frame() { --nframes; }
  std::string key_str(Key k) const { return ac.sym.str(k); }

  // ---- contexts (runtime.rs:71-117, 151-187)
  void push_scope() { Frame& f = ctx(); f.marks.push_back((uint32_t)f.size()); }
  void pop_scope() {
    Frame& f = ctx();
    const size_t mark = f.marks.back();
    f.marks.pop_back();
    // :180-184 forced merge: the return variable is carried into the enclosing scope
    size_t keep = f.size();
    if (!f.marks.empty())
      for (size_t i = mark; i < f.size(); ++i)
        if (f.keys[i] == kReturn) { keep = i; break; }
    if (keep < f.size()) {
      Item carried = std::move(f.items[keep]);
      f.truncate(mark);
      f.push(kReturn, std::move(carried));
    } else f.truncate(mark);
  }
  Item& declare_item(DataType type, Key name, const Path& dims) {  // runtime.rs:190-222
    Frame& f = ctx();
    Item* existing = f.find(name);
    if (existing && type != D_Variable) runtime_error("Item already declared");
    Item item;
    item.type = type;
    if (type == D_Signal) {
      if (dims.empty()) item.sig.val = next_signal_id++;
      else item.sig = nested_new<uint32_t>(dims, 0, [&] { return next_signal_id++; });  // row-major ids, :431-445
    } else if (type == D_Variable) {
      if (!dims.empty()) item.var = nested_new<std::optional<uint32_t>>(dims, 0, [] { return std::optional<uint32_t>(); });
    } else if (!dims.empty()) item.comp = nested_new<SigMap>(dims, 0, [] { return SigMap(); });
    if (existing) { *existing = std::move(item); return *existing; }  // a re-declared variable overwrites the visible one (:175-178)
    return f.push(name, std::move(item));
  }
  Item& item(Key name, const char* who) {
    Item* it = ctx().find(name);
    if (!it) runtime_error(std::string("Item not declared: ") + who + ": " + key_str(name));
    return *it;
  }
  DataType type_of(const Ref& r) {  // runtime.rs:235-249
    switch (r.kind) {
      case Ref::TempVar: return D_Variable;
      case Ref::TempSignal: return D_Signal;
      case Ref::TempComp: return D_Component;
      default: return item(r.name, "get_item_data_type").type;
    }
  }
  std::string ref_name(const Ref& r) const { return r.kind == Ref::Named ? key_str(r.name) : std::string(); }
  std::optional<uint32_t> variable_value(const Ref& r) {  // runtime.rs:296-307
    if (r.kind == Ref::TempVar) return r.value;
    if (r.kind != Ref::Named) runtime_error("Item not declared: get_variable_value: " + ref_name(r));
    Item& it = item(r.name, "get_variable_value");
    if (it.type != D_Variable) runtime_error("Item not declared: get_variable_value: " + ref_name(r));
    if (r.access.empty() && !it.var.is_array) return it.var.val;
    const auto& n = nested_get(it.var, access_to_u32(r.access));
    if (n.is_array) runtime_error("Data Item content is not a single value");
    return n.val;
  }
  uint32_t value_or_empty(const Ref& r) {
    auto v = variable_value(r);
    if (!v) fail(C2A_PROG_EMPTY_DATA_ITEM, "Empty data item");
    return *v;
  }
  // (component access, signal access) split, runtime.rs:617-663
  static void split_component_access(const Ref& r, Path& comp_path, Key& signal, Path& sig_path) {
    bool has = false;
    for (uint32_t k = 0; k < r.access.size(); ++k) {
      const SubAccess& s = r.access[k];
      if (!s.component) (has ? sig_path : comp_path).push_back((uint32_t)s.v);
      else { if (has) runtime_error("Access Error"); signal = s.v; has = true; }
    }
    if (!has) runtime_error("Access Error");
  }
  const SigTree& component_signal_content(const Ref& r) {  // runtime.rs:389-405, 560-580
    Path cp, sp;
    Key sname = 0;
    split_component_access(r, cp, sname, sp);
    const SigMap* map;
    if (r.kind == Ref::TempComp) { if (!cp.empty()) runtime_error("Access Error"); map = r.comp.get(); }
    else {
      Item& it = item(r.name, "get_component_signal_id");
      if (it.type != D_Component) runtime_error("Item not declared: get_component_signal_id: " + ref_name(r));
      const auto& n = nested_get(it.comp, cp);
      if (n.is_array) runtime_error("Data Item content is not a single value");
      map = &n.val;
    }
    const SigTree* f = map->find(sname);
    if (!f) runtime_error("Item not declared: get_signal_id: " + key_str(sname));
    return nested_get(*f, sp);
  }
  const SigTree& signal_content(const Ref& r) {  // runtime.rs:323-337
    static thread_local SigTree tmp;
    if (r.kind == Ref::TempSignal) { tmp.is_array = false; tmp.arr.clear(); tmp.val = r.signal; return tmp; }
    Item& it = item(r.name, "get_signal_content");
    if (it.type != D_Signal) runtime_error("Item not declared: get_signal_content: " + ref_name(r));
    if (r.access.empty()) return it.sig;
    return nested_get(it.sig, access_to_u32(r.access));
  }
  uint32_t signal_id(const Ref& r) {  // runtime.rs:341-355
    const SigTree& n = signal_content(r);
    if (n.is_array) runtime_error("Data Item content is not a single value");
    return n.val;
  }
  uint32_t component_signal_id(const Ref& r) {
    const SigTree& n = component_signal_content(r);
    if (n.is_array) runtime_error("Data Item content is not a single value");
    return n.val;
  }
  const SigTree& content_for_access(const Ref& r) {  // process.rs:582-591
    switch (type_of(r)) {
      case D_Signal: return signal_content(r);
      case D_Component: return component_signal_content(r);
      default: fail(C2A_PROG_INVALID_DATA_TYPE, "Invalid data type");
    }
  }

  // ---- process.rs
  uint32_t make_constant(uint32_t value) {  // :558-579 — one constant signal per value per visible context
    const Key name = kConstKey | value;
    Item* it = ctx().find(name);
    if (it && it->type == D_Signal && !it->sig.is_array) return it->sig.val;
    uint32_t id = declare_item(D_Signal, name, Path()).sig.val;
    ac.add_signal(id, ctx().ctx, name, nullptr, false, value);
    return id;
  }
  uint32_t signal_for_access(const Ref& r) {  // :538-556
    switch (type_of(r)) {
      case D_Signal: return signal_id(r);
      case D_Variable: return make_constant(value_or_empty(r));
      default: return component_signal_id(r);
    }
  }
  Ref temp_signal() {  // declare_random_item(Signal) + add_signal, :464-474
    Ref r;
    r.kind = Ref::TempSignal;
    r.signal = next_signal_id++;
    ac.add_signal(r.signal, ctx().ctx, 0, nullptr, true, std::nullopt);
    return r;
  }
  static uint32_t execute_op(uint32_t l, uint32_t r, int op) {  // :649-750 (release-build wrapping for + * ** << >>)
    auto op_err = [](const char* m) { fail(C2A_PROG_OPERATION_ERROR, std::string("Operation error: ") + m); };
    switch (op) {
      case I_Mul: return l * r;
      case I_Div: if (!r) op_err("Division by zero"); return l / r;
      case I_Add: return l + r;
      case I_Sub: if (l < r) op_err("Subtraction underflow"); return l - r;
      case I_Pow: { uint32_t acc = 1, base = l, e = r; while (e) { if (e & 1) acc *= base; e >>= 1; base *= base; } return acc; }
      case I_IntDiv: if (!r) op_err("Integer division by zero"); return l / r;
      case I_Mod: if (!r) op_err("Modulo by zero"); return l % r;
      case I_ShiftL: return l << (r & 31);
      case I_ShiftR: return l >> (r & 31);
      case I_LesserEq: return l <= r;
      case I_GreaterEq: return l >= r;
      case I_Lesser: return l < r;
      case I_Greater: return l > r;
      case I_Eq: return l == r;
      case I_NotEq: return l != r;
      case I_BoolOr: return l != 0 || r != 0;
      case I_BoolAnd: return l != 0 && r != 0;
      case I_BitOr: return l | r;
      case I_BitAnd: return l & r;
      default: return l ^ r;
    }
  }
  Ref temp_var(std::optional<uint32_t> v) { Ref r; r.kind = Ref::TempVar; r.value = v; return r; }

  Ref build_access(Key name, const std::vector<Access>& access) {  // :620-646
    Ref r;
    r.kind = Ref::Named;
    r.name = name;
    for (auto& a : access) {
      if (a.component) r.access.push_back(SubAccess{true, a.key});
      else r.access.push_back(SubAccess{false, value_or_empty(process_expression(*a.index))});
    }
    return r;
  }

  Ref process_expression(const Expr& e) {  // :280-312
    switch (e.kind) {
      case Expr::Call: return handle_call(e);
      case Expr::InfixOp: {  // :426-478
        Ref l = process_expression(*e.l);
        Ref r = process_expression(*e.r);
        DataType lt = type_of(l), rt = type_of(r);
        if (lt == D_Variable && rt == D_Variable) return temp_var(execute_op(value_or_empty(l), value_or_empty(r), e.op));
        uint32_t lid = signal_for_access(l);
        uint32_t rid = signal_for_access(r);
        Ref out = temp_signal();
        ac.add_gate(kGateOf[e.op], lid, rid, out.signal);
        return out;
      }
      case Expr::PrefixOp: {  // :485-533, 758-764
        Ref r = process_expression(*e.r);
        uint32_t lhs_value = e.op == P_Complement ? 0xFFFFFFFFu : 0u;
        int infix = e.op == P_Sub ? I_Sub : (e.op == P_BoolNot ? I_Eq : I_BitXor);
        if (type_of(r) == D_Variable) return temp_var(execute_op(lhs_value, value_or_empty(r), infix));
        uint32_t lid = make_constant(lhs_value);
        uint32_t rid = signal_for_access(r);
        Ref out = temp_signal();
        ac.add_gate(kGateOf[infix], lid, rid, out.signal);
        return out;
      }
      case Expr::Number:  // :294-306: the literal must fit u32
        if (e.num_overflow) fail(C2A_PROG_PARSING_ERROR, "Parsing error");
        return temp_var(e.num);
      case Expr::Variable:
        if (e.access.empty()) {  // a plain variable is read here: the same value every later look-up of the name would give
          Item* it = ctx().find(e.key);  // (the callee of a call in between runs in its own frame and cannot write it)
          if (it && it->type == D_Variable && !it->var.is_array) return temp_var(it->var.val);
        }
        return build_access(e.key, e.access);
      default: fail(C2A_PROG_EXPRESSION_NOT_IMPLEMENTED, "Expression not implemented");
    }
  }

  // ---- instance memo.  A call runs in a fresh, empty context (runtime.rs:75-77): it sees its argument values and nothing else,
  // so what it emits is a function of (callable, arguments) up to the signal id it starts at - every id in its calls is one it
  // allocated itself (its own signals, constants and temporaries, and those of the calls it makes).  From the second instance
  // on, the walk of a (callable, arguments) pair is therefore replayed from the first one's slice of the recorded stream with
  // the ids shifted, instead of being interpreted again (the reference interprets every instance).  Kept per pair: the
  // slice, the id range, the result (component signal map / function value) and the call depth it needed (:call error).
  struct MemoKey {
    const Callable* c;
    std::vector<uint32_t> args;
    bool operator==(const MemoKey& o) const { return c == o.c && args == o.args; }
  };
  struct MemoHash {
    size_t operator()(const MemoKey& k) const {
      uint64_t h = (uint64_t)(uintptr_t)k.c * 0x9E3779B97F4A7C15ull;
      for (uint32_t a : k.args) h = (h ^ a) * 0x100000001B3ull + (h >> 29);
      return (size_t)h;
    }
  };
  struct Memo {
    uint32_t seen = 0;
    bool cached = false;
    Sink::Mark begin{}, end{};  // the instance's slice of the recording
    uint32_t gen = 0;           // largest replay generation inside the slice
    uint64_t rel_depth = 0;  // deepest call below this one, relative to it
    std::optional<uint32_t> value;
    std::shared_ptr<SigMap> comp;
  };
  std::unordered_map<MemoKey, Memo, MemoHash> memo;
  bool memo_on = true;
  uint64_t depth_hw = 0;  // deepest call since the enclosing call began
  uint64_t memo_hits = 0;

  static void shift_ids(SigTree& t, uint32_t delta) {
    if (!t.is_array) { t.val += delta; return; }
    for (auto& x : t.arr) shift_ids(x, delta);
  }

  Ref handle_call(const Expr& e) {  // :315-419
    if (!e.callee) fail(C2A_PROG_UNDEFINED_CALLABLE, "Undefined function or template");
    const Callable& c = *e.callee;
    Small<uint32_t, 8> args;
    for (auto& a : e.args) args.push_back(value_or_empty(process_expression(*a)));
    if (++depth > 10000) fail(C2A_PROG_CALL_ERROR, "Call error");
    Memo* m = nullptr;
    bool record = false;
    if (memo_on) {
      MemoKey key{&c, {}};
      key.args.reserve(args.size());
      for (uint32_t i = 0; i < args.size(); ++i) key.args.push_back(args[i]);
      m = &memo[std::move(key)];  // (references into an unordered_map survive the insertions the nested calls make)
      if (m->cached && depth + m->rel_depth <= 10000) {
        const uint32_t delta = next_signal_id - m->begin.id;
        ac.replay(m->begin, m->end, delta, m->gen);
        next_signal_id += m->end.id - m->begin.id;
        depth_hw = std::max(depth_hw, depth + m->rel_depth);
        ++memo_hits;
        Ref hit;
        if (c.is_function) { hit.kind = Ref::TempVar; hit.value = m->value; }
        else {
          hit.kind = Ref::TempComp;
          hit.comp = std::make_shared<SigMap>(*m->comp);
          for (auto& kv : hit.comp->v) shift_ids(kv.second, delta);
        }
        --depth;
        return hit;
      }
      record = !m->cached && m->seen++ >= 1;  // the second instance is the one that is kept: a pair seen once costs nothing
    }
    const Sink::Mark begin = ac.mark();
    const uint64_t saved_hw = depth_hw;
    depth_hw = depth;
    push_frame((uint32_t)e.key);  // push_context(false, id): empty context named after the callee
    for (size_t i = 0; i < c.param_keys.size() && i < args.size(); ++i)
      declare_item(D_Variable, c.param_keys[i], Path()).var.val = args[(uint32_t)i];
    process_statements(c.body);
    Ref ret;
    if (c.is_function) {
      ret.kind = Ref::TempVar;
      Item* it = ctx().find(kReturn);
      if (it && it->type == D_Variable && !it->var.is_array) ret.value = it->var.val;
    } else {
      ret.kind = Ref::TempComp;
      ret.comp = std::make_shared<SigMap>();
      ret.comp->v.reserve(c.input_keys.size() + c.output_keys.size());
      for (auto* list : {&c.input_keys, &c.output_keys})
        for (Key name : *list) {
          Item* it = ctx().find(name);
          if (!it || it->type != D_Signal) runtime_error("Item not declared: get_signal: " + key_str(name));
          ret.comp->set(name, it->sig);
        }
    }
    pop_frame();
    if (record) {
      m->begin = begin; m->end = ac.mark();
      m->gen = ac.max_gen_between(m->begin, m->end);
      m->rel_depth = depth_hw - depth;
      m->value = ret.value;
      if (ret.comp) m->comp = std::make_shared<SigMap>(*ret.comp);
      m->cached = true;
    }
    depth_hw = std::max(depth_hw, saved_hw);
    --depth;
    return ret;
  }

  void connect_signal_arrays(const SigTree& a, const SigTree& b) {  // :594-617
    if (!a.is_array || !b.is_array || a.arr.size() != b.arr.size()) fail(C2A_PROG_INVALID_DATA_TYPE, "Invalid data type");
    for (size_t i = 0; i < a.arr.size(); ++i) {
      if (!a.arr[i].is_array && !b.arr[i].is_array) ac.add_connection(a.arr[i].val, b.arr[i].val);
      else if (a.arr[i].is_array && b.arr[i].is_array) connect_signal_arrays(a.arr[i], b.arr[i]);
      else fail(C2A_PROG_INVALID_DATA_TYPE, "Invalid data type");
    }
  }

  // NOTE: an Item& / SigTree& into the frame dies with the next declare_item (make_constant declares): what is needed
  // across such a call is copied first (scalars are the common case and copy without touching the heap).
  void handle_substitution(const Stmt& s) {  // :192-277
    Ref lh = build_access(s.key, s.access);
    Ref rh = process_expression(*s.e);
    const DataType target = item(s.key, "get_item_data_type").type;
    switch (target) {
      case D_Variable: {
        std::optional<uint32_t> v = variable_value(rh);
        Item& it = item(s.key, "set_variable");
        auto& slot = lh.access.empty() ? it.var : nested_get_mut(it.var, access_to_u32(lh.access));
        if (slot.is_array) runtime_error("Data Item content is not a single value");
        slot.val = v;
        break;
      }
      case D_Component:
        if (s.op == A_Var) {  // component instantiation
          SigMap map;
          if (rh.kind == Ref::TempComp) {
            if (rh.comp.use_count() == 1) map = std::move(*rh.comp);
            else map = *rh.comp;
          } else {
            if (rh.kind != Ref::Named) runtime_error("Item not declared: get_component_map: " + ref_name(rh));
            Item& src = item(rh.name, "get_component_map");
            if (src.type != D_Component) runtime_error("Item not declared: get_component_map: " + ref_name(rh));
            const auto& n = nested_get(src.comp, access_to_u32(rh.access));
            if (n.is_array) runtime_error("Data Item content is not a single value");
            map = n.val;
          }
          auto& slot = nested_get_mut(item(s.key, "set_component").comp, access_to_u32(lh.access));
          if (slot.is_array) runtime_error("Data Item content is not a single value");
          slot.val = std::move(map);
        } else if (s.op == A_ConstraintSignal) {
          const SigTree& lhs_ref = component_signal_content(lh);
          if (lhs_ref.is_array) {  // (nothing is declared while two arrays are wired: the references stay valid)
            const SigTree& rhs_content = content_for_access(rh);
            if (!rhs_content.is_array) fail(C2A_PROG_INVALID_DATA_TYPE, "Invalid data type");
            connect_signal_arrays(lhs_ref, rhs_content);
          } else {
            uint32_t comp_sig = lhs_ref.val;
            uint32_t assigned = signal_for_access(rh);
            ac.add_connection(assigned, comp_sig);
          }
        } else fail(C2A_PROG_OPERATION_NOT_SUPPORTED, "Operation not supported");
        break;
      case D_Signal:
        if (s.e->kind == Expr::Variable) {
          const SigTree& lhs_ref = signal_content(lh);
          if (lhs_ref.is_array) {
            const SigTree& rhs_content = content_for_access(rh);
            if (!rhs_content.is_array) fail(C2A_PROG_INVALID_DATA_TYPE, "Invalid data type");
            connect_signal_arrays(lhs_ref, rhs_content);
          } else {
            uint32_t lhs_id = lhs_ref.val;
            uint32_t out = signal_for_access(rh);
            ac.add_connection(out, lhs_id);
          }
        } else if (s.e->kind == Expr::Call || s.e->kind == Expr::InfixOp || s.e->kind == Expr::PrefixOp || s.e->kind == Expr::Number) {
          uint32_t given = signal_id(lh);
          uint32_t out = signal_for_access(rh);
          ac.add_connection(out, given);
        } else fail(C2A_PROG_SIGNAL_SUBSTITUTION_NOT_IMPLEMENTED, "Signal substitution not implemented");
        break;
    }
  }

  void process_statements(const std::vector<StmtP>& v) { for (auto& s : v) process_statement(*s); }

  void announce_signals(const SigTree& n, Key name, Path& at) {  // one add_signal per element, row-major (:79-108)
    if (!n.is_array) { ac.add_signal(n.val, ctx().ctx, name, &at, false, std::nullopt); return; }
    for (uint32_t i = 0; i < n.arr.size(); ++i) { at.push_back(i); announce_signals(n.arr[i], name, at); at.pop_back(); }
  }

  void process_statement(const Stmt& s) {  // :36-189
    switch (s.kind) {
      case Stmt::InitBlock:
      case Stmt::Block: process_statements(s.stmts); break;
      case Stmt::Substitution: handle_substitution(s); break;
      case Stmt::Declaration: {
        Path dims;
        if (!s.dims.empty()) {
          std::vector<Ref> dim_refs;
          for (auto& d : s.dims) dim_refs.push_back(process_expression(*d));
          for (auto& r : dim_refs) dims.push_back(value_or_empty(r));
        }
        Item& it = declare_item(s.dtype, s.key, dims);
        if (s.dtype == D_Signal) {
          // a zero-length dimension: the reference's index loop still visits index 0 once and fails (:96, IndexOutOfBounds)
          for (uint32_t k = 0; k < dims.size(); ++k) if (dims[k] == 0) runtime_error("Index out of bounds");
          Path at;
          announce_signals(it.sig, s.key, at);
        }
        break;
      }
      case Stmt::IfThenElse: {
        uint32_t c = value_or_empty(process_expression(*s.e));
        const Stmt* branch = c ? s.a.get() : s.b.get();
        if (branch) { push_scope(); process_statement(*branch); pop_scope(); }
        break;
      }
      case Stmt::While: {
        push_scope();  // WHILE_PRE
        while (true) {
          uint32_t c = value_or_empty(process_expression(*s.e));
          if (!c) break;
          push_scope();  // WHILE_EXE
          process_statement(*s.a);
          pop_scope();
        }
        pop_scope();
        break;
      }
      case Stmt::Return: {  // :160-174 — no early exit in the reference either
        uint32_t v = value_or_empty(process_expression(*s.e));
        declare_item(D_Variable, kReturn, Path()).var.val = v;
        break;
      }
      case Stmt::Assert:
        if (!value_or_empty(process_expression(*s.e))) runtime_error("Assertion failed");
        break;
      default: fail(C2A_PROG_STATEMENT_NOT_IMPLEMENTED, "Statement not implemented");
    }
  }
};

}  // namespace front

// ============================================================================================================ C ABI
struct c2a_program {
  front::Sink sink;
  std::string error;
  std::vector<uint32_t> inputs, outputs;          // signal ids tagged by the prefix match of src/program.rs:57-66, ascending
  std::vector<std::string> main_inputs, main_outputs;  // declared names of the main template
  std::unordered_map<uint32_t, std::string> spelled;   // names handed out by c2a_program_signal_name (pointers stay valid)
};

static int compile_body(c2a_program* p, const std::string& src, const std::string& file, const std::string& dir, c2a_compiler* into) {
  using namespace front;
  p->sink = Sink();
  p->sink.into = into;
  p->sink.lazy = into == nullptr;
  p->error.clear();
  p->spelled.clear();
  p->inputs.clear(); p->outputs.clear();
  try {
    Program prog;
    std::set<std::string> seen;
    parse_into(prog, src, file, dir, seen);
    if (!prog.main || prog.main->kind != Expr::Call) fail(C2A_PROG_MAIN_NOT_A_CALL, "Main expression not a call");  // program.rs:68
    p->sink.sym.annotate(prog);
    const Callable* maindef = prog.main->callee;
    if (!maindef || maindef->is_function) fail(C2A_PROG_UNDEFINED_CALLABLE, "Undefined function or template");
    const Callable& main = *maindef;
    Walker w(prog, p->sink);
    if (const char* off = getenv("C2A_FRONT_NO_MEMO")) w.memo_on = !(off[0] && off[0] != '0');  // diagnostic: interpret every instance
    std::vector<std::optional<uint32_t>> values;  // program.rs:30-37
    for (auto& a : prog.main->args) values.push_back(w.variable_value(w.process_expression(*a)));
    for (size_t i = 0; i < main.param_keys.size() && i < values.size(); ++i)  // :40-51 declared in the ROOT context
      w.declare_item(D_Variable, main.param_keys[i], Path()).var.val = values[i];
    w.process_statements(main.body);  // :54-55
    p->main_inputs = main.inputs;
    p->main_outputs = main.outputs;
    // :57-66 — prefix match over ALL signal names ("0.c" also tags "0.const_signal_*"): replicate.  Only names of the root
    // context can start with "0." (every other context is named after a template or function).
    const uint32_t root = p->sink.sym.intern("0");
    // A name is "0." + base + "[i][j]..": an identifier key is a prefix of it iff it is a prefix of the base, so the test is made
    // once per base symbol and no name is spelled unless a host emitter wants it.
    auto tag = [&](const std::vector<std::string>& keys, std::vector<uint32_t>& ids, bool input) {
      auto starts = [&](const std::string& base) {
        for (auto& k : keys) if (base.compare(0, k.size(), k) == 0) return true;
        return false;
      };
      auto may_match = [&](const char* stem) {  // some key agrees with the stem on their common length ("c" and "const_signal_")
        const size_t L = strlen(stem);
        for (auto& k : keys) if (k.compare(0, std::min(L, k.size()), stem, std::min(L, k.size())) == 0) return true;
        return false;
      };
      const bool const_may = may_match("const_signal_"), random_may = may_match("random_");
      std::vector<int8_t> by_sym(p->sink.sym.strs.size(), -1);
      size_t sp = 0;  // replayed id ranges belong to callee contexts (and their name records are not materialised): skip them
      for (uint32_t id = 0; id < p->sink.names.size(); ++id) {
        while (sp < p->sink.spans.size() && p->sink.spans[sp].dst + p->sink.spans[sp].len <= id) ++sp;
        if (sp < p->sink.spans.size() && p->sink.spans[sp].dst <= id) { id = p->sink.spans[sp].dst + p->sink.spans[sp].len - 1; continue; }
        const SigName& n = p->sink.names[id];
        if (n.ctx != root) continue;
        bool hit;
        switch (n.kind_n & 3u) {
          case 0: {
            int8_t& c = by_sym[n.a];
            if (c < 0) c = starts(p->sink.sym.strs[n.a]) ? 1 : 0;
            hit = c == 1;
            break;
          }
          case 1: hit = const_may && starts("const_signal_" + std::to_string(n.a)); break;
          default: hit = random_may && starts("random_" + std::to_string(id)); break;
        }
        if (!hit) continue;
        ids.push_back(id);  // ascending ids = the deterministic stand-in for the reference's HashMap order
        if (into) (input ? c2a_add_input : c2a_add_output)(into, id, p->sink.name_of(id).c_str());
      }
    };
    tag(main.inputs, p->inputs, true);
    tag(main.outputs, p->outputs, false);
  } catch (const front::Error& e) {
    p->error = e.text;
    return e.code;
  } catch (const std::bad_alloc&) {
    p->error = "out of memory";
    return C2A_ERR_NO_MEMORY;
  }
  return C2A_OK;
}

// The walk recurses once per nested call / statement / expression level; the call-depth guard (10 000 nested calls, "Call error")
// must be reached before the native stack ends, so the walk runs on a thread of its own with a 512 MB stack (address space only:
// pages are touched as the recursion deepens).  The reference has no guard at all (a runaway recursion aborts the process).
struct CompileJob {
  c2a_program* p;
  const std::string *src, *file, *dir;
  c2a_compiler* into;
  int status;
};
static void* compile_thread(void* arg) {
  CompileJob* j = (CompileJob*)arg;
  j->status = compile_body(j->p, *j->src, *j->file, *j->dir, j->into);
  return nullptr;
}
static int compile_impl(c2a_program* p, const std::string& src, const std::string& file, const std::string& dir, c2a_compiler* into) {
  CompileJob job{p, &src, &file, &dir, into, C2A_OK};
  pthread_attr_t attr;
  pthread_t th;
  bool threaded = false;
  if (pthread_attr_init(&attr) == 0) {
    if (pthread_attr_setstacksize(&attr, (size_t)512 << 20) == 0 && pthread_create(&th, &attr, compile_thread, &job) == 0) threaded = true;
    pthread_attr_destroy(&attr);
  }
  if (!threaded) return compile_body(p, src, file, dir, into);  // (no address space for the big stack: the caller's stack has to do)
  pthread_join(th, nullptr);
  return job.status;
}

extern "C" {

c2a_program* c2a_program_new(void) { return new c2a_program(); }
void c2a_program_free(c2a_program* p) { delete p; }
const char* c2a_program_error(const c2a_program* p) { return p ? p->error.c_str() : "null program"; }

int c2a_program_compile_file(c2a_program* p, const char* path, c2a_compiler* into) {
  if (!p || !path) return C2A_ERR_INVALID_ARGUMENT;
  std::string src;
  try { src = front::read_file(path); } catch (const front::Error& e) { p->error = e.text; return e.code; }
  return compile_impl(p, src, path, front::dir_of(path), into);
}
int c2a_program_compile_source(c2a_program* p, const char* source, const char* include_dir, c2a_compiler* into) {
  if (!p || !source) return C2A_ERR_INVALID_ARGUMENT;
  return compile_impl(p, source, "<source>", include_dir ? include_dir : ".", into);
}
int c2a_program_packed(c2a_program* p, c2a_packed_events* out) {  // the recording itself: nothing is converted
  if (!p || !out) return C2A_ERR_INVALID_ARGUMENT;
  p->sink.materialise();
  *out = c2a_packed_events{p->sink.kinds.data(), p->sink.words.data(), (uint64_t)p->sink.kinds.size(), (uint64_t)p->sink.words.size(),
                           C2A_PACKED_DENSE_IDS, 0};
  return C2A_OK;
}
int c2a_program_compressed(c2a_program* p, c2a_compressed_events* out) {  // nothing is expanded here
  if (!p || !out) return C2A_ERR_INVALID_ARGUMENT;
  *out = c2a_compressed_events{p->sink.kinds.data(), p->sink.words.data(), (uint64_t)p->sink.kinds.size(), (uint64_t)p->sink.words.size(),
                               p->sink.replays.data(), (uint64_t)p->sink.replays.size(), p->sink.max_gen, C2A_PACKED_DENSE_IDS};
  return C2A_OK;
}
uint64_t c2a_program_num_events(const c2a_program* p) { return p->sink.kinds.size(); }
const c2a_event* c2a_program_events(const c2a_program* p) { return const_cast<c2a_program*>(p)->sink.events().data(); }
uint64_t c2a_program_num_constants(const c2a_program* p) { return p->sink.const_ids.size(); }
const uint32_t* c2a_program_constant_signals(const c2a_program* p) { return p->sink.const_ids.data(); }
const uint32_t* c2a_program_constant_values(const c2a_program* p) { return p->sink.const_vals.data(); }
uint64_t c2a_program_num_signals(const c2a_program* p) { return p->sink.names.size(); }
const char* c2a_program_signal_name(const c2a_program* cp, uint32_t id) {  // spelled on first request, then kept
  if (!cp || id >= cp->sink.names.size()) return nullptr;
  c2a_program* p = const_cast<c2a_program*>(cp);
  auto it = p->spelled.find(id);
  if (it == p->spelled.end()) it = p->spelled.emplace(id, p->sink.name_of(id)).first;
  return it->second.c_str();
}
// many names at once: '\n'-separated (a signal id that does not exist contributes an empty name); returns the byte count
uint64_t c2a_program_signal_names(const c2a_program* p, const uint32_t* ids, uint64_t n, char* out, uint64_t cap) {
  if (!p || (n && !ids)) return 0;
  uint64_t at = 0;
  for (uint64_t i = 0; i < n; ++i) {
    std::string nm = ids[i] < p->sink.names.size() ? p->sink.name_of(ids[i]) : std::string();
    if (out && at + nm.size() + 1 <= cap) { memcpy(out + at, nm.data(), nm.size()); out[at + nm.size()] = '\n'; }
    at += nm.size() + 1;
  }
  return at;
}
uint32_t c2a_program_num_inputs(const c2a_program* p) { return (uint32_t)p->inputs.size(); }
uint32_t c2a_program_num_outputs(const c2a_program* p) { return (uint32_t)p->outputs.size(); }
const uint32_t* c2a_program_inputs(const c2a_program* p) { return p->inputs.data(); }
const uint32_t* c2a_program_outputs(const c2a_program* p) { return p->outputs.data(); }

}  // extern "C"
