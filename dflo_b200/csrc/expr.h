// Expressions in x,y,t as dflo writes them in input.prm (deal.II FunctionParser = muparser,
// reference src/parameters.cc:470-526): compiled once on the host to a small
// postfix program and evaluated by the same DFLO_HD interpreter on the host (initial
// conditions) and on the device (time-dependent boundary values at every RK stage, e.g. the
// moving-shock top wall of examples/double_mach_reflection/input.prm:35-41).
//
// Grammar (lowest to highest precedence): || ; && ; == != ; < <= > >= ; + - ; * / ; unary - ; ^
// (right associative); functions sin cos tan asin acos atan sinh cosh tanh exp log ln log10 sqrt
// abs sign rint min max pow atan2 if; constants _pi _e; variables x y t (z accepted, = 0).
// Comparisons and logic evaluate to 1.0 / 0.0 as in muparser.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define DFLO_EXPR_HD __host__ __device__
#else
#define DFLO_EXPR_HD
#endif

namespace dflo
{
   enum ExprOp
   {
      OP_CONST = 0, OP_X, OP_Y, OP_T, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_NEG,
      OP_LT, OP_LE, OP_GT, OP_GE, OP_EQ, OP_NE, OP_AND, OP_OR,
      OP_SIN, OP_COS, OP_TAN, OP_ASIN, OP_ACOS, OP_ATAN, OP_SINH, OP_COSH, OP_TANH,
      OP_EXP, OP_LOG, OP_LOG10, OP_SQRT, OP_ABS, OP_SIGN, OP_RINT, OP_MIN, OP_MAX, OP_ATAN2, OP_IF
   };

   struct ExprInstr
   {
      int op;
      int pad;
      double val;
   };

   constexpr int EXPR_STACK = 24;

   DFLO_EXPR_HD inline double expr_eval (const ExprInstr *code, int n, double x, double y, double t)
   {
      double st[EXPR_STACK];
      int sp = 0;
      for (int i = 0; i < n; ++i)
      {
         const int op = code[i].op;
         switch (op)
         {
            case OP_CONST: st[sp++] = code[i].val; break;
            case OP_X: st[sp++] = x; break;
            case OP_Y: st[sp++] = y; break;
            case OP_T: st[sp++] = t; break;
            case OP_NEG: st[sp - 1] = -st[sp - 1]; break;
            case OP_SIN: st[sp - 1] = sin (st[sp - 1]); break;
            case OP_COS: st[sp - 1] = cos (st[sp - 1]); break;
            case OP_TAN: st[sp - 1] = tan (st[sp - 1]); break;
            case OP_ASIN: st[sp - 1] = asin (st[sp - 1]); break;
            case OP_ACOS: st[sp - 1] = acos (st[sp - 1]); break;
            case OP_ATAN: st[sp - 1] = atan (st[sp - 1]); break;
            case OP_SINH: st[sp - 1] = sinh (st[sp - 1]); break;
            case OP_COSH: st[sp - 1] = cosh (st[sp - 1]); break;
            case OP_TANH: st[sp - 1] = tanh (st[sp - 1]); break;
            case OP_EXP: st[sp - 1] = exp (st[sp - 1]); break;
            case OP_LOG: st[sp - 1] = log (st[sp - 1]); break;
            case OP_LOG10: st[sp - 1] = log10 (st[sp - 1]); break;
            case OP_SQRT: st[sp - 1] = sqrt (st[sp - 1]); break;
            case OP_ABS: st[sp - 1] = fabs (st[sp - 1]); break;
            case OP_SIGN: st[sp - 1] = (st[sp - 1] > 0) ? 1.0 : (st[sp - 1] < 0 ? -1.0 : 0.0); break;
            case OP_RINT: st[sp - 1] = rint (st[sp - 1]); break;
            case OP_IF:
               sp -= 2;
               st[sp - 1] = (st[sp - 1] != 0.0) ? st[sp] : st[sp + 1];
               break;
            default:
            {
               const double b = st[--sp];
               const double a = st[sp - 1];
               double r = 0.0;
               switch (op)
               {
                  case OP_ADD: r = a + b; break;
                  case OP_SUB: r = a - b; break;
                  case OP_MUL: r = a * b; break;
                  case OP_DIV: r = a / b; break;
                  case OP_POW: r = pow (a, b); break;
                  case OP_LT: r = a < b ? 1.0 : 0.0; break;
                  case OP_LE: r = a <= b ? 1.0 : 0.0; break;
                  case OP_GT: r = a > b ? 1.0 : 0.0; break;
                  case OP_GE: r = a >= b ? 1.0 : 0.0; break;
                  case OP_EQ: r = a == b ? 1.0 : 0.0; break;
                  case OP_NE: r = a != b ? 1.0 : 0.0; break;
                  case OP_AND: r = (a != 0.0 && b != 0.0) ? 1.0 : 0.0; break;
                  case OP_OR: r = (a != 0.0 || b != 0.0) ? 1.0 : 0.0; break;
                  case OP_MIN: r = a < b ? a : b; break;
                  case OP_MAX: r = a > b ? a : b; break;
                  case OP_ATAN2: r = atan2 (a, b); break;
                  default: break;
               }
               st[sp - 1] = r;
            }
         }
      }
      return sp > 0 ? st[sp - 1] : 0.0;
   }
}

#include <cctype>
#include <cstdlib>
#include <string>
#include <vector>

namespace dflo
{
   // Recursive-descent compiler to postfix.  Returns false and sets err on a syntax error.
   class ExprCompiler
   {
   public:
      bool compile (const std::string &text, std::vector<ExprInstr> &out, std::string &err, bool *uses_t = nullptr)
      {
         s = text;
         pos = 0;
         code.clear ();
         error.clear ();
         depth = max_depth = 0;
         has_t = false;
         parse_or ();
         skip ();
         if (error.empty () && pos != s.size ()) error = "unexpected '" + s.substr (pos, 1) + "'";
         if (error.empty () && max_depth > EXPR_STACK) error = "expression too deep";
         if (!error.empty ())
         {
            err = error + " in \"" + text + "\"";
            return false;
         }
         out = code;
         if (uses_t) *uses_t = has_t;
         return true;
      }

   private:
      std::string s, error;
      size_t pos;
      std::vector<ExprInstr> code;
      int depth, max_depth;
      bool has_t;

      void skip ()
      {
         while (pos < s.size () && std::isspace ((unsigned char) s[pos])) ++pos;
      }
      bool eat (const char *tok)
      {
         skip ();
         size_t n = 0;
         while (tok[n]) ++n;
         if (s.compare (pos, n, tok) == 0)
         {
            pos += n;
            return true;
         }
         return false;
      }
      void emit (int op, double v = 0.0, int delta = 0)
      {
         ExprInstr i;
         i.op = op;
         i.pad = 0;
         i.val = v;
         code.push_back (i);
         depth += delta;
         if (depth > max_depth) max_depth = depth;
      }
      void binary (int op) { emit (op, 0.0, -1); }

      void parse_or ()
      {
         parse_and ();
         while (error.empty () && eat ("||"))
         {
            parse_and ();
            binary (OP_OR);
         }
      }
      void parse_and ()
      {
         parse_eq ();
         while (error.empty () && eat ("&&"))
         {
            parse_eq ();
            binary (OP_AND);
         }
      }
      void parse_eq ()
      {
         parse_rel ();
         while (error.empty ())
         {
            if (eat ("=="))
            {
               parse_rel ();
               binary (OP_EQ);
            }
            else if (eat ("!="))
            {
               parse_rel ();
               binary (OP_NE);
            }
            else
               break;
         }
      }
      void parse_rel ()
      {
         parse_add ();
         while (error.empty ())
         {
            if (eat ("<="))
            {
               parse_add ();
               binary (OP_LE);
            }
            else if (eat (">="))
            {
               parse_add ();
               binary (OP_GE);
            }
            else if (eat ("<"))
            {
               parse_add ();
               binary (OP_LT);
            }
            else if (eat (">"))
            {
               parse_add ();
               binary (OP_GT);
            }
            else
               break;
         }
      }
      void parse_add ()
      {
         parse_mul ();
         while (error.empty ())
         {
            if (eat ("+"))
            {
               parse_mul ();
               binary (OP_ADD);
            }
            else if (eat ("-"))
            {
               parse_mul ();
               binary (OP_SUB);
            }
            else
               break;
         }
      }
      void parse_mul ()
      {
         parse_unary ();
         while (error.empty ())
         {
            if (eat ("*"))
            {
               parse_unary ();
               binary (OP_MUL);
            }
            else if (eat ("/"))
            {
               parse_unary ();
               binary (OP_DIV);
            }
            else
               break;
         }
      }
      void parse_unary ()
      {
         if (eat ("-"))
         {
            parse_unary ();
            emit (OP_NEG);
         }
         else if (eat ("+"))
            parse_unary ();
         else
            parse_pow ();
      }
      void parse_pow ()
      {
         parse_primary ();
         if (error.empty () && eat ("^"))
         {
            parse_unary (); // right associative, binds tighter than unary minus on its left
            binary (OP_POW);
         }
      }
      void parse_primary ()
      {
         skip ();
         if (pos >= s.size ())
         {
            error = "unexpected end of expression";
            return;
         }
         const char ch = s[pos];
         if (ch == '(')
         {
            ++pos;
            parse_or ();
            if (!eat (")")) error = "missing ')'";
            return;
         }
         if (std::isdigit ((unsigned char) ch) || ch == '.')
         {
            char *end = nullptr;
            const double v = std::strtod (s.c_str () + pos, &end);
            pos = end - s.c_str ();
            emit (OP_CONST, v, +1);
            return;
         }
         if (std::isalpha ((unsigned char) ch) || ch == '_')
         {
            size_t e = pos;
            while (e < s.size () && (std::isalnum ((unsigned char) s[e]) || s[e] == '_')) ++e;
            const std::string name = s.substr (pos, e - pos);
            pos = e;
            if (name == "x") { emit (OP_X, 0, +1); return; }
            if (name == "y") { emit (OP_Y, 0, +1); return; }
            if (name == "z") { emit (OP_CONST, 0.0, +1); return; }
            if (name == "t") { has_t = true; emit (OP_T, 0, +1); return; }
            if (name == "_pi" || name == "pi") { emit (OP_CONST, 3.14159265358979323846, +1); return; }
            if (name == "_e") { emit (OP_CONST, 2.71828182845904523536, +1); return; }
            static const struct { const char *n; int op; int nargs; } fn[] = {
               {"sin", OP_SIN, 1}, {"cos", OP_COS, 1}, {"tan", OP_TAN, 1}, {"asin", OP_ASIN, 1}, {"acos", OP_ACOS, 1},
               {"atan", OP_ATAN, 1}, {"sinh", OP_SINH, 1}, {"cosh", OP_COSH, 1}, {"tanh", OP_TANH, 1}, {"exp", OP_EXP, 1},
               {"log", OP_LOG, 1}, {"ln", OP_LOG, 1}, {"log10", OP_LOG10, 1}, {"sqrt", OP_SQRT, 1}, {"abs", OP_ABS, 1},
               {"sign", OP_SIGN, 1}, {"rint", OP_RINT, 1}, {"min", OP_MIN, 2}, {"max", OP_MAX, 2}, {"pow", OP_POW, 2},
               {"atan2", OP_ATAN2, 2}, {"if", OP_IF, 3}};
            for (auto &f : fn)
               if (name == f.n)
               {
                  if (!eat ("("))
                  {
                     error = "missing '(' after " + name;
                     return;
                  }
                  for (int a = 0; a < f.nargs; ++a)
                  {
                     parse_or ();
                     if (!error.empty ()) return;
                     if (a + 1 < f.nargs && !eat (","))
                     {
                        error = "missing ',' in " + name;
                        return;
                     }
                  }
                  if (!eat (")"))
                  {
                     error = "missing ')' after " + name;
                     return;
                  }
                  emit (f.op, 0.0, 1 - f.nargs);
                  return;
               }
            error = "unknown identifier '" + name + "'";
            return;
         }
         error = std::string ("unexpected '") + ch + "'";
      }
   };
}
