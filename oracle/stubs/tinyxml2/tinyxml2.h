// Small DOM stand-in for tinyxml2 (pinned by the reference at 9a89766, not in this image).
// Test infrastructure for oracle/_ref only: it lets the reference's own MJCF reader
// (src/xml/*.cc) parse the model files. Covers the subset of the tinyxml2 API those
// files name: elements, attributes, comments, deep clone, child insertion/removal and a
// plain printer. Text nodes, declarations and DTDs are parsed and dropped.
#ifndef ORACLE_STUB_TINYXML2_H_
#define ORACLE_STUB_TINYXML2_H_

#include <cstddef>
#include <string>

namespace tinyxml2 {

class XMLDocument;
class XMLElement;
class XMLComment;
class XMLPrinter;

enum XMLError { XML_SUCCESS = 0, XML_ERROR_PARSING = 1 };

class XMLAttribute {
 public:
  const char* Name() const { return name_.c_str(); }
  const char* Value() const { return value_.c_str(); }
  const XMLAttribute* Next() const { return next_; }

 private:
  friend class XMLElement;
  friend class XMLDocument;
  std::string name_, value_;
  XMLAttribute* next_ = nullptr;
};

class XMLNode {
 public:
  virtual ~XMLNode();
  virtual XMLElement* ToElement() { return nullptr; }
  virtual const XMLElement* ToElement() const { return nullptr; }
  virtual XMLComment* ToComment() { return nullptr; }
  virtual const XMLComment* ToComment() const { return nullptr; }
  virtual XMLNode* ShallowClone(XMLDocument* doc) const = 0;

  const char* Value() const { return value_.c_str(); }
  void SetValue(const char* v) { value_ = v; }
  int GetLineNum() const { return line_; }
  XMLDocument* GetDocument() { return doc_; }
  const XMLDocument* GetDocument() const { return doc_; }
  XMLNode* Parent() { return parent_; }
  const XMLNode* Parent() const { return parent_; }
  bool NoChildren() const { return first_ == nullptr; }
  XMLNode* FirstChild() { return first_; }
  const XMLNode* FirstChild() const { return first_; }
  XMLNode* LastChild() { return last_; }
  XMLNode* NextSibling() { return next_; }
  const XMLNode* NextSibling() const { return next_; }
  XMLNode* PreviousSibling() { return prev_; }

  XMLElement* FirstChildElement(const char* name = nullptr);
  const XMLElement* FirstChildElement(const char* name = nullptr) const;
  XMLElement* NextSiblingElement(const char* name = nullptr);
  const XMLElement* NextSiblingElement(const char* name = nullptr) const;

  XMLNode* InsertEndChild(XMLNode* add);
  XMLNode* LinkEndChild(XMLNode* add) { return InsertEndChild(add); }
  XMLNode* InsertFirstChild(XMLNode* add);
  XMLNode* InsertAfterChild(XMLNode* after, XMLNode* add);
  void DeleteChild(XMLNode* node);
  void DeleteChildren();
  XMLNode* DeepClone(XMLDocument* target) const;

 protected:
  explicit XMLNode(XMLDocument* doc) : doc_(doc) {}
  friend class XMLDocument;
  friend class XMLPrinter;
  void Unlink(XMLNode* child);
  XMLDocument* doc_;
  XMLNode* parent_ = nullptr;
  XMLNode* first_ = nullptr;
  XMLNode* last_ = nullptr;
  XMLNode* prev_ = nullptr;
  XMLNode* next_ = nullptr;
  std::string value_;
  int line_ = 0;
};

class XMLComment : public XMLNode {
 public:
  XMLComment* ToComment() override { return this; }
  const XMLComment* ToComment() const override { return this; }
  XMLNode* ShallowClone(XMLDocument* doc) const override;

 private:
  friend class XMLDocument;
  explicit XMLComment(XMLDocument* doc) : XMLNode(doc) {}
};

class XMLElement : public XMLNode {
 public:
  ~XMLElement() override;
  XMLElement* ToElement() override { return this; }
  const XMLElement* ToElement() const override { return this; }
  XMLNode* ShallowClone(XMLDocument* doc) const override;

  const char* Name() const { return Value(); }
  void SetName(const char* n) { SetValue(n); }
  const char* Attribute(const char* name, const char* value = nullptr) const;
  const XMLAttribute* FirstAttribute() const { return attr_; }
  const XMLAttribute* FindAttribute(const char* name) const;
  void SetAttribute(const char* name, const char* value);
  void SetAttribute(const char* name, int value);
  void SetAttribute(const char* name, double value);
  void DeleteAttribute(const char* name);
  const char* GetText() const { return nullptr; }

 private:
  friend class XMLDocument;
  friend class XMLPrinter;
  explicit XMLElement(XMLDocument* doc) : XMLNode(doc) {}
  XMLAttribute* attr_ = nullptr;
};

class XMLPrinter {
 public:
  explicit XMLPrinter(void* file = nullptr, bool compact = false, int depth = 0)
      : compact_(compact) { (void)file; (void)depth; }
  virtual ~XMLPrinter() {}
  const char* CStr() const { return buf_.c_str(); }
  int CStrSize() const { return (int)buf_.size() + 1; }
  virtual void PrintSpace(int depth) { for (int i = 0; i < depth; ++i) Write("    "); }
  void Write(const char* s) { buf_ += s; }
  void PrintNode(const XMLNode* node, int depth);

 private:
  std::string buf_;
  bool compact_;
};

class XMLDocument : public XMLNode {
 public:
  XMLDocument() : XMLNode(nullptr) { doc_ = this; }
  ~XMLDocument() override {}
  XMLNode* ShallowClone(XMLDocument*) const override { return nullptr; }

  XMLError Parse(const char* xml, size_t nbytes = static_cast<size_t>(-1));
  bool Error() const { return err_ != XML_SUCCESS; }
  XMLError ErrorID() const { return err_; }
  const char* ErrorStr() const { return errstr_.c_str(); }
  int ErrorLineNum() const { return errline_; }
  void ClearError() { err_ = XML_SUCCESS; errstr_.clear(); }
  XMLElement* RootElement() { return FirstChildElement(); }
  const XMLElement* RootElement() const { return FirstChildElement(); }
  XMLElement* NewElement(const char* name);
  XMLComment* NewComment(const char* text);
  void Print(XMLPrinter* printer = nullptr) const;
  void Clear() { DeleteChildren(); ClearError(); }

 private:
  void Fail(int line, const std::string& msg);
  XMLError err_ = XML_SUCCESS;
  std::string errstr_;
  int errline_ = 0;
};

}  // namespace tinyxml2

#endif  // ORACLE_STUB_TINYXML2_H_
